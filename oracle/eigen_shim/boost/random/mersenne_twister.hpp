// Stand-in for Boost.Random 1.84 (conanfile.py:54), absent from this image.  Only the names
// slam.h:23-36,759 needs.  The variates differ from Boost's (different normal algorithm), which is
// why the draws are INPUTS everywhere else in this repo.
#pragma once
#include <random>
namespace boost {
using mt19937 = std::mt19937;
}
