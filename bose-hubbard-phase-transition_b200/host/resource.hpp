// resource.hpp -- drop-in for the reference's namespace Resource (include/resource.hpp:16-38), Eigen-free part.
// The OpenMP thread heuristic (set_omp_threads, src/resource.cpp:77-83) has no counterpart: grid points are
// scheduled over GPUs, not host threads.
#pragma once

#include <chrono>

namespace Resource {
long get_memory_usage(bool print = false);  // resident set size in KB (/proc/self/statm)
long get_available_memory();                // free RAM in KB
void timer();                               // first call starts, second call prints "Calculation duration: ..."
}  // namespace Resource
