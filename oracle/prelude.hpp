// Pre-included when the reference's examples/LaplaceBEM.cpp is compiled UNCHANGED for oracle/_ref:
// the example relies on Boost pulling these in, and uses an unqualified isnan
// (reference examples/BEM/Triangulation.hpp:196).  TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include <deque>
#include <string>
#include <iostream>
using std::isnan;
