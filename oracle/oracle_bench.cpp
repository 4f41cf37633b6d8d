// oracle/oracle_bench.cpp — TEST/BENCH INFRASTRUCTURE: CPU baseline timing loops (filled in below).
#include "slam_oracle.hpp"
extern "C" int orc_bench_placeholder(void) { return 0; }
