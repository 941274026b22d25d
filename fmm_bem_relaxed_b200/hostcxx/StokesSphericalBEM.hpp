#pragma once
/** @file StokesSphericalBEM.hpp
 * The reference's examples/LaplaceBEM.cpp includes this header first (it brings in Mat3.hpp for the mesh generators)
 * without using the kernel class.  The StokesSphericalBEM kernel itself (reference kernel/StokesSphericalBEM.hpp) is
 * not built on the GPU yet; this header only provides what that driver needs from it.
 */
#include "Mat3.hpp"
