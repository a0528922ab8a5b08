// sx_host.h -- host-side objects shared by the translation units of the library (not part of the C ABI).
#pragma once

#include <map>
#include <mutex>
#include <string>

#include "../../include/stratego_b200.h"
#include "sx_device.cuh"

struct LaunchPlan {
    int warps_per_block, blocks_per_sm, smem_per_block, num_sms, grid, regs;
    int warp_bytes, tile_bytes;
};

struct sx_config {
    sx::DevConfig dev;
    sx_layout layout;
    int cells_per_lane;  // K
    int games_per_warp;  // G
    bool compact_movers;  // gen_moves<.., COMPACT>: dense 10x10 boards (sx_device.cuh)
    bool compact_forced = false;  // sx_config_set_tuning chose compact_movers: it then applies to every launch mode
    sx_state start_states{nullptr, nullptr, nullptr};  // sx_config_set_start_states: curriculum table (device)
    long long n_start_states = 0;
    int32_t *start_index = nullptr;  // [env id - start_index_base] entry each game was started from (optional)
    long long start_index_base = 0, start_index_len = 0;
    int tune_warps = 0;   // sx_config_set_tuning: resident warps per SM of the warp-level kernel (0 = built-in choice)
    int tune_issue = -1;  // sx_config_set_tuning: where a game's background copy is issued (-1 = built-in choice)
    // launch shapes already worked out, keyed by (device, ops, mode): the occupancy / attribute queries cost tens
    // of microseconds, which matters for the one-game API
    mutable std::mutex plan_mutex;
    mutable std::map<uint64_t, LaunchPlan> plans;
};

int sx_set_error(const std::string &msg);  // records the text sx_last_error() returns; always returns -1
