// Pre-included when a reference driver is compiled UNCHANGED against these headers: the reference relies on Boost
// pulling in a few standard headers and uses an unqualified isnan (examples/BEM/Triangulation.hpp:196,
// examples/BEM/MeshIO.hpp:13).
#include <cmath>
#include <cstring>
#include <sstream>
using std::isnan;
