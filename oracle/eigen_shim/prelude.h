// oracle/eigen_shim/prelude.h — force-included (-include) when compiling the reference's sources
// with GCC: slam.h uses the MSVC-internal `std::_Pi_val` (a double-valued pi constant), e.g.
// slam/include/slam.h:66-67,73,81,818-825.
#pragma once
namespace std {
inline constexpr double _Pi_val = 3.14159265358979323846;
}
