// Global count of this library's kernel launches (reported by czk_ctx_launches / bench.py's gpu_launches).
#pragma once
#include <atomic>
#include <cstdint>
namespace czk {
extern std::atomic<uint64_t> g_launch_count;
}
#define CZK_LAUNCHED() (::czk::g_launch_count.fetch_add(1, std::memory_order_relaxed))
