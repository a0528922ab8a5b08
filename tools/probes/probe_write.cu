// HBM write-ceiling probe (development aid): how fast can 8 GB be written with different store paths?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_st_v4(uint4 *p, size_t n) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_st_cs(uint4 *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
__global__ void k_st_evict_first(uint4 *p, size_t n) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4), "l"(pol) : "memory");
}
// contiguous chunk per block (each block streams its own region) instead of grid-interleaved
__global__ void k_st_chunk(uint4 *p, size_t n) {
    const size_t per = n / gridDim.x;
    uint4 *q = p + per * blockIdx.x;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) q[i] = v;
}
// TMA bulk stores: each warp owns a `tile`-byte shared buffer and streams it to consecutive tiles
template <bool EVICT_FIRST>
__global__ void k_tma(uint8_t *p, size_t n_tiles, int tile) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    uint8_t *buf = smem + (size_t)warp * tile;
    for (int i = lane * 16; i < tile; i += 512) *reinterpret_cast<uint4 *>(buf + i) = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint64_t pol = 0;
    if (EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (lane == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
        for (size_t t = blockIdx.x * (size_t)wpb + warp; t < n_tiles; t += (size_t)gridDim.x * wpb) {
            if (EVICT_FIRST)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p + t * tile), "r"(s), "r"(tile), "l"(pol) : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * tile), "r"(s), "r"(tile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
// the engine's output layout: per game a 26 800 B observation (16 B aligned) and a 3 700 B mask row (4 B aligned),
// written with the same TMA bulk + head/tail word stores, nothing else (no logic, no sparse entries)
// state_mode: 0 none; 1 per-game ~176 B state reads (scattered among the output stream); 2 per-game reads + write-back;
// 3 reads from a small L2-resident buffer (what a resident state would cost); 4 contiguous 32-game chunks per warp:
// one coalesced 5.6 KB state read per chunk; 5 = 4 + coalesced write-back per chunk; 6 contiguous chunks, no state
__global__ void k_layout(uint8_t *obs, uint8_t *mask, size_t n_envs, int obs_bytes, int mask_bytes, int wait_read, uint32_t *state = nullptr, uint32_t *sink = nullptr, int state_mode = 1) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = threadIdx.x * 16; i < obs_bytes + mask_bytes + 32; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t s_obs = (uint32_t)__cvta_generic_to_shared(smem), s_mask = s_obs + ((obs_bytes + 15) & ~15);
    uint32_t acc = 0;
    auto emit = [&](size_t e) {
        uint8_t *go = obs + e * obs_bytes, *gm = mask + e * mask_bytes;
        const int head = int((16 - (reinterpret_cast<uintptr_t>(gm) & 15)) & 15), body = (mask_bytes - head) & ~15, tail = mask_bytes - head - body;
        if (lane < (head >> 2)) reinterpret_cast<uint32_t *>(gm)[lane] = 0;
        if (lane >= 8 && lane - 8 < (tail >> 2)) *reinterpret_cast<uint32_t *>(gm + head + body + ((lane - 8) << 2)) = 0;
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(go), "r"(s_obs), "r"(obs_bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gm + head), "r"(s_mask), "r"(body) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (wait_read) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    };
    if (state && state_mode >= 4) {
        const size_t n_chunks = n_envs / 32;
        for (size_t c = blockIdx.x * (size_t)wpb + warp; c < n_chunks; c += (size_t)gridDim.x * wpb) {
            uint32_t v[40];
            if (state_mode != 6) {
#pragma unroll
                for (int j = 0; j < 40; ++j) v[j] = state[c * 1280 + j * 32 + lane];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc += state[n_envs * 40 + c * 128 + j * 32 + lane];
#pragma unroll
                for (int j = 0; j < 40; ++j) acc += v[j];
            }
            for (int i = 0; i < 32; ++i) emit(c * 32 + i);
            if (state_mode == 5) {
#pragma unroll
                for (int j = 0; j < 40; ++j) state[c * 1280 + j * 32 + lane] = v[j] + 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) state[n_envs * 40 + c * 128 + j * 32 + lane] = acc;
            }
        }
    } else {
        for (size_t e = blockIdx.x * (size_t)wpb + warp; e < n_envs; e += (size_t)gridDim.x * wpb) {
            if (state) {
                const size_t se = state_mode == 3 ? (e & 4095) : e;
                const uint32_t a = state[se * 40 + lane], b = state[n_envs * 40 + se * 4 + (lane & 3)];  // ~176 B of per-game state reads
                acc += a + b;
                if (state_mode == 2) { state[se * 40 + lane] = a + 1; if (lane < 4) state[n_envs * 40 + se * 4 + lane] = b + 1; }
            }
            emit(e);
        }
    }
    if (sink && acc == 0x12345678u) sink[0] = acc;
}
__global__ void k_copy(const uint4 *a, uint4 *b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
    const size_t bytes = 8ull << 30;
    uint8_t *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *name, double nbytes, int reps) {
        float ms; cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.3f ms  %7.0f GB/s\n", name, ms / reps, nbytes * reps / ms / 1e6);
    };
    const int reps = 5;
    const size_t n16 = bytes / 16;
#define RUN(name, nbytes, launch) do { launch; launch; cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < reps; ++r) { launch; } cudaEventRecord(e1); CK(cudaGetLastError()); report(name, nbytes, reps); } while (0)
    RUN("cudaMemsetAsync", (double)bytes, cudaMemsetAsync(a, 1, bytes));
    for (int bps : {1, 2, 4, 8, 16}) {
        char nm[64]; snprintf(nm, 64, "st.v4 grid=148x%d block=256", bps);
        RUN(nm, (double)bytes, (k_st_v4<<<148 * bps, 256>>>((uint4 *)a, n16)));
    }
    RUN("st.v4 grid=148x8 block=1024", (double)bytes, (k_st_v4<<<148 * 2, 1024>>>((uint4 *)a, n16)));
    RUN("st.cs.v4 grid=148x8 block=256", (double)bytes, (k_st_cs<<<148 * 8, 256>>>((uint4 *)a, n16)));
    RUN("st.evict_first.v4 grid=148x8 block=256", (double)bytes, (k_st_evict_first<<<148 * 8, 256>>>((uint4 *)a, n16)));
    RUN("st.v4 contiguous chunk/block 148x8", (double)bytes, (k_st_chunk<<<148 * 8, 256>>>((uint4 *)a, n16)));
    for (int tile : {4096, 16384, 30720}) {
        for (int wpb : {4, 7}) {
            if ((size_t)wpb * tile > 227 * 1024) continue;
            cudaFuncSetAttribute(k_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wpb * tile);
            cudaFuncSetAttribute(k_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wpb * tile);
            char nm[64]; snprintf(nm, 64, "TMA bulk tile=%d warps/SM=%d", tile, wpb);
            RUN(nm, (double)(bytes / tile) * tile, (k_tma<false><<<148, wpb * 32, wpb * tile>>>(a, bytes / tile, tile)));
            snprintf(nm, 64, "TMA bulk evict_first tile=%d warps/SM=%d", tile, wpb);
            RUN(nm, (double)(bytes / tile) * tile, (k_tma<true><<<148, wpb * 32, wpb * tile>>>(a, bytes / tile, tile)));
        }
    }
    {
        const int ob = 26800, mb = 3700;
        const size_t n_envs = 262144;
        uint8_t *mask = b;
        cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
        for (int wpb : {8, 10, 12, 16}) for (int wr : {1, 0}) {
            char nm[96]; snprintf(nm, 96, "engine layout obs26800+mask3700 W=%d wait=%s", wpb, wr ? "read" : "full");
            RUN(nm, (double)n_envs * (ob + mb), (k_layout<<<148, wpb * 32, 32768>>>(a, mask, n_envs, ob, mb, wr)));
            if (!wr) {
                const char *names[7] = {"", "+ per-game state reads", "+ per-game state reads + write-back", "+ state reads from an L2-resident buffer",
                                        "contiguous 32-game chunks, coalesced state read", "contiguous chunks, coalesced read + write-back", "contiguous chunks, no state"};
                for (int sm : {1, 2, 3, 4, 5, 6}) {
                    snprintf(nm, 96, "  %-52s W=%d", names[sm], wpb);
                    RUN(nm, (double)n_envs * (ob + mb), (k_layout<<<148, wpb * 32, 32768>>>(a, mask, n_envs, ob, mb, wr, (uint32_t *)(b + (4ull << 30)), (uint32_t *)(b + (6ull << 30)), sm)));
                }
            }
        }
    }
    RUN("copy ld.v4/st.v4 grid=148x8", 2.0 * bytes, (k_copy<<<148 * 8, 256>>>((const uint4 *)a, (uint4 *)b, n16)));
    RUN("cudaMemcpyAsync D2D", 2.0 * bytes, cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
    return 0;
}
