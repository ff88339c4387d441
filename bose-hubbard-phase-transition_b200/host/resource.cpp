#include "resource.hpp"

#include <sys/sysinfo.h>
#include <unistd.h>

#include <fstream>
#include <iostream>

namespace {
std::chrono::high_resolution_clock::time_point g_start;
bool g_running = false;
}  // namespace

long Resource::get_memory_usage(bool print)
{
    long kb = -1;
    std::ifstream f("/proc/self/statm");
    long size = 0, resident = 0;
    if (f >> size >> resident) kb = resident * (sysconf(_SC_PAGESIZE) / 1024);
    if (kb < 0)
        std::cerr << "Error reading memory usage from /proc/self/statm." << std::endl;
    else if (print)
        std::cout << "Memory usage: " << kb << " KB" << std::endl;
    return kb;
}

long Resource::get_available_memory()
{
    struct sysinfo info;
    if (sysinfo(&info) != 0) {
        std::cerr << "Error getting system info." << std::endl;
        return -1;
    }
    return info.freeram / 1024;
}

void Resource::timer()
{
    if (!g_running) {
        g_start = std::chrono::high_resolution_clock::now();
        g_running = true;
        return;
    }
    const std::chrono::duration<double> d = std::chrono::high_resolution_clock::now() - g_start;
    g_running = false;
    if (d.count() > 60) {
        const int minutes = static_cast<int>(d.count()) / 60;
        std::cout << "Calculation duration: " << minutes << " minutes " << d.count() - minutes * 60 << " seconds." << std::endl;
    } else {
        std::cout << "Calculation duration: " << d.count() << " seconds." << std::endl;
    }
}
