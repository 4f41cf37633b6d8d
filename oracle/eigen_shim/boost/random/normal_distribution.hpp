#pragma once
#include <random>
namespace boost {
template <class T = double>
using normal_distribution = std::normal_distribution<T>;
}
