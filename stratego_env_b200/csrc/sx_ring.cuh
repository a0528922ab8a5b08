// sx_ring.cuh -- output rendering of the fused step through a per-warp ring of shared-memory chunks.
//
// The general kernel (sx_kernels.cu, sx_fused_kernel<..., RING = 0>) writes a game's outputs as "background image by TMA
// bulk copy + sparse 4-byte stores to global memory": cheap in instructions, but every sparse store is a partial-sector
// write that L2 has to merge into a line the copy engine wrote a moment ago.  Barrage pays ~60 of them per game, Standard
// ~280, and that is exactly the order of their roofline fractions (0.86 / 0.79).
//
// With RING > 0 nothing sparse ever reaches global memory.  A game's observation is cut into chunks of CHUNK_CELLS board
// cells (20 cells x 67 floats = 5 360 bytes; any multiple of 4 cells is a multiple of 16 bytes).  Each warp owns RING
// chunk slots: it copies the (cell-periodic) background into a slot, adds the chunk's state-dependent entries with plain
// shared-memory stores, and hands the slot to the TMA engine (cp.async.bulk shared -> global).  Before a slot is reused
// the warp waits only until the copy that used it has been READ out of shared memory (cp.async.bulk.wait_group.read), never
// for the write to land, so a warp renders chunk i + 1 while chunk i drains and goes on to its next game's rules while
// the last chunks are still in flight.  The mask row is rendered the same way in a slot of its own and leaves in the
// first chunk's bulk group.  Every byte of output is written exactly once, as full lines, by the copy engine.
#pragma once

#include "sx_device.cuh"

namespace sx {
namespace ring {

#ifndef SX_CHUNK_CELLS
#define SX_CHUNK_CELLS 20
#endif
constexpr int CHUNK_CELLS = SX_CHUNK_CELLS;  // cells per chunk: a multiple of 4 (16-byte granularity of bulk copies), at most 32 (one cell per lane)
static_assert(CHUNK_CELLS % 4 == 0 && CHUNK_CELLS >= 4 && CHUNK_CELLS <= 32, "chunk = 4..32 cells in multiples of 4");

__host__ __device__ inline int chunk_bytes(int channels) { return CHUNK_CELLS * channels * 4; }

// per-warp ring: [mask slot][RING x (partial-observation chunk | full-observation chunk)]
__host__ __device__ inline int mask_slot_bytes(const DevConfig &cfg, bool mask) { return mask ? round16(cfg.mask_bytes + 16) : 0; }
__host__ __device__ inline int slot_bytes(const DevConfig &cfg, bool po, bool fo)
{
    return (po ? chunk_bytes(cfg.po_ch) : 0) + (fo ? chunk_bytes(cfg.fo_ch) : 0);
}
__host__ __device__ inline int warp_ring_bytes(const DevConfig &cfg, bool po, bool fo, bool mask, int slots)
{
    return mask_slot_bytes(cfg, mask) + slots * slot_bytes(cfg, po, fo);
}
// the block's background images: CHUNK_CELLS cells of each observation (the background is periodic in the cell)
__host__ __device__ inline int background_bytes(const DevConfig &cfg, bool po, bool fo) { return slot_bytes(cfg, po, fo); }

template <int SLOTS>
__device__ __forceinline__ void wait_slot_free()
{
    // at most SLOTS - 1 bulk groups may still be reading shared memory: the group that used the slot about to be
    // rewritten (SLOTS groups ago) is not one of them
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SLOTS - 1) : "memory");
}

__device__ __forceinline__ void copy16(uint8_t *dst, const uint8_t *src, int bytes, int lane)
{
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const int n = bytes >> 4;
#pragma unroll 4
    for (int i = lane; i < n; i += 32) d[i] = s[i];
}

// The state-dependent entries of cells [c0, c0 + n) of observer `me`'s observation (impl:1232-1397 + maenv:499-508),
// stored on top of the background held in `slot` (shared memory; slot[q * channels + ch] is cell c0 + q).
// Same entries as patch_obs of the general kernel, one cell per lane.
__device__ __forceinline__ void patch_chunk(const DevConfig &cfg, const WarpMem &m, const Aux &a, float *slot, const ObsMap om,
                                            int me, int c0, int n)
{
    const int lane = lane_id(), flip = me, CH = om.channels;
    const float one = cfg.unit_lut[1];
    if (lane < n) {
        const uint32_t b = m.board[view(c0 + lane, flip, cfg.N)];
        float *cell = slot + lane * CH;
        if (b & CELL_OBST) cell[om.obstacle] = one;
        const int rank = b & CELL_RANK;
        if (rank) {
            const int po = (b & CELL_REVEALED) ? rank : SP_UNKNOWN;
            const bool own = int((b >> 4) & 1) == me;
            if (own) {
                cell[om.own_true + rank - 1] = one;
                cell[om.own_po + po - 1] = one;
                if (b & CELL_STILL) cell[om.own_still] = one;
            } else {
                if (om.enemy_true >= 0) cell[om.enemy_true + rank - 1] = one;
                cell[om.enemy_po + po - 1] = one;
                if (b & CELL_STILL) cell[om.enemy_still] = one;
            }
        }
    }
    if (lane < 4) {  // recent-move squares: lanes 0/1 own from/to, lanes 2/3 enemy from/to
        const int who = (lane < 2) ? me : (me ^ 1);
        const int cell_abs = (lane & 1) ? a.rto[who] : a.rfrom[who];
        const int code = (lane & 1) ? -a.rcode[who] : 1;
        const unsigned q = unsigned(view(cell_abs, flip, cfg.N) - c0);
        if (cell_abs != NO_CELL && q < unsigned(n)) slot[q * CH + (lane < 2 ? om.own_recent : om.enemy_recent)] = cfg.recent_lut[code + 3];
    }
    for (int e = lane; e < a.ncap; e += 32) {
        const uint32_t ent = m.cap[e];
        const int cell_abs = ent & 0xff, owner = (ent >> 8) & 1, type0 = (ent >> 9) & 15, count = int(ent >> 13) + 1;
        const unsigned q = unsigned(view(cell_abs, flip, cfg.N) - c0);
        if (q < unsigned(n)) slot[q * CH + (owner == me ? om.own_cap : om.enemy_cap) + type0] = cfg.cap_lut[type0 * 9 + count];
    }
}

}  // namespace ring
}  // namespace sx
