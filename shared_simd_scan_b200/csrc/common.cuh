// common.cuh -- device-side building blocks shared by every kernel family (sm_100a only).
//
//   * mbarrier + cp.async.bulk (TMA 1-D bulk copy) wrappers,
//   * TileRing: a multi-stage shared-memory ring fed by one elected thread with bulk copies of
//     contiguous packed-column tiles (persistent CTAs, static round-robin tile order),
//   * warp/CTA reduction of match counts into a 64-bit global counter.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/simd_scan_commons.hpp"

#if !defined(__CUDA_ARCH__) || (__CUDA_ARCH__ >= 1000)
#else
#error "shared_simd_scan_b200 kernels are written for sm_100a only"
#endif

namespace sss {

constexpr int kThreads = 256;  // every streaming kernel uses 256-thread CTAs
constexpr int kWarps = kThreads / 32;

// ---- PTX wrappers ----------------------------------------------------------------------------

// Programmatic dependent launch (kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization):
// pdl_enter() at the very top of a kernel lets the NEXT kernel of the stream be scheduled as this one's CTAs
// leave, and then waits until the PREVIOUS kernel has completed and its writes are visible.  Nothing before
// it may touch global memory.  Launched without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter()
{
    pdl_trigger();
    pdl_wait();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// make mbarrier.init visible to the async (TMA) proxy before the first bulk copy targets it
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// order prior generic-proxy smem reads/writes before subsequent async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// L2 eviction policy for data that is read exactly once (the packed column)
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// the same without a cache hint (short columns that a caller scans again and again should stay in L2)
__device__ __forceinline__ void bulk_g2s_plain(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 256-bit store (sm_100: STG.E.256): one full 32-byte sector per lane.  p must be 32-byte aligned.
__device__ __forceinline__ void st_v8_u32(void* p, const uint32_t (&w)[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// streaming stores: results are written once and never re-read by these kernels
// (-DSSS_PLAIN_STORES=1: developer switch, plain stores everywhere these helpers are used)
#ifndef SSS_PLAIN_STORES
#define SSS_PLAIN_STORES 0
#endif
#if SSS_PLAIN_STORES
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) { *p = v; }
__device__ __forceinline__ void st_stream_v4(int4* p, int4 v) { *p = v; }
#else
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_v4(int4* p, int4 v) { __stcs(p, v); }
#endif

// ---- tile ring --------------------------------------------------------------------------------
//
// The packed column is cut into tiles of kTileBytes contiguous bytes (a whole number of 32-code
// blocks, so every tile starts 16-byte aligned for any w).  Persistent CTAs draw tiles from a ticket
// counter (below); kStages bulk copies are kept in flight; all threads wait on the stage's "full"
// mbarrier, consume the tile from shared memory, and whoever releases the stage last refills it.
//
// Shared memory layout: [kStages][kTileBytes + kPadBytes], kStages mbarriers, kStages tile numbers.
// Ring depth.  Measured on B200 (profiles/r01_ring_depth_ab.txt): with two CTAs per SM the scan
// peaks with about 64 KB of ring per CTA (~100 KB of loads in flight per SM); deeper rings are
// SLOWER (104 KB per CTA costs 5-10 % at W = 8..17), shallower ones starve.  2..8 stages.
#ifndef SSS_RING_KB
#define SSS_RING_KB 64
#endif
constexpr int ring_stages(int tile_bytes)
{
    const int fit = (SSS_RING_KB * 1024) / (tile_bytes + 16);
    return fit < 2 ? 2 : (fit > 8 ? 8 : fit);
}

// Tile order.  Round 1 gave CTA b the tiles b, b + grid, b + 2 grid, ... (static round-robin).  A
// persistent copy kernel written that way reaches 5.9-6.2 TB/s on a B200, the same kernel drawing its
// tiles from an atomic ticket counter 6.85 TB/s (tools/probes/copy_patterns.cu,
// profiles/r02_copy_patterns.txt): SMs do not all stream at the same speed (two dies, far and near L2
// slices), and with a static split the launch lasts as long as its slowest SM.  So the ring draws its
// tiles: whoever refills a stage takes the next ticket, writes the tile number next to the stage and
// issues the bulk copy; consumers learn the tile from there.  A ticket past the last tile is recorded
// as kNoTile and completes the stage's barrier without a copy -- the CTA's cue to leave.
// `tickets` = {next tile, CTAs finished}: context-owned, zero between launches, re-armed by the last
// CTA to finish (finish()).
constexpr uint64_t kNoTile = ~0ull;

template <int kTileBytes, int kStages, int kPadBytes = 16, bool kDynamic = true>
struct TileRing {
    static_assert(kTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    static constexpr int kStageBytes = kTileBytes + kPadBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStages * 8 + kStages * 8 + kStages * 8 + 16;
    static_assert(kStageBytes % 16 == 0, "stage base must stay 16-byte aligned");

    uint8_t* base;
    uint64_t* full;
    uint64_t* tile_of;    // per stage: the tile it holds (or will hold), kNoTile past the end
    // per stage (release_warp protocol): low kCountBits = warps that have finished reading it, the bits above =
    // the ticket for its refill, delivered by the arrival of the lane that drew it (26 bits: 64 M tiles; the
    // launchers' columns stay far below -- a tile is at least 8 KB)
    static constexpr int kCountBits = 6;
    static constexpr uint32_t kTicketMask = (1u << (32 - kCountBits)) - 1u;
    uint32_t* released;
    static __device__ __forceinline__ uint32_t clamp_ticket(uint64_t t) { return t < kTicketMask ? (uint32_t)t : kTicketMask; }
    uint64_t in_hand, in_hand2;  // the same for release(): registers of thread 0
    // release_warp(): tickets this lane drew and will hand on with its next even / odd arrival (kNoTicket:
    // none; tickets fit 32 bits: a tile is >= 2 KB)
    static constexpr uint32_t kNoTicket = 0xFFFFFFFFu;
    uint32_t pend_even, pend_odd;
    const uint8_t* src;
    uint64_t src_bytes;   // readable bytes of the packed column, rounded up to 16
    uint64_t num_tiles;
    uint64_t policy;
    unsigned long long* tickets;

    __device__ __forceinline__ TileRing(uint8_t* smem, const void* packed, uint64_t readable_bytes16,
                                        unsigned long long* tickets_)
        : base(smem), full(reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes)),
          tile_of(reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kStages * 8)),
          released(reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + kStages * 16)), in_hand(0), in_hand2(0),
          pend_even(kNoTicket), pend_odd(kNoTicket),
          src(static_cast<const uint8_t*>(packed)), src_bytes(readable_bytes16),
          num_tiles((readable_bytes16 + kTileBytes - 1) / kTileBytes), policy(0), tickets(tickets_)
    {
    }

    __device__ __forceinline__ const uint32_t* stage_words(int stage) const
    {
        return reinterpret_cast<const uint32_t*>(base + stage * kStageBytes);
    }

    // A ticket costs a round trip to L2 (~0.4 us), too much to sit between "stage free" and "copy issued"
    // when a tile is worth a microsecond of work or less (11-13 KB scan tiles: 0.98 -> 0.83 of the peak
    // with the draw inline; the K <= 8 shared scans -10..20 %).  So tickets are drawn AHEAD and nobody ever
    // waits for one:
    //   release()       thread 0 keeps the next two tickets in registers; the one it draws now is used two
    //                   tiles later
    //   release_warp()  the refilling lane draws the next ticket of its parity into a register and hands it
    //                   on with ITS arrival two tiles later (in the high half of the stage's arrival
    //                   counter), where the last reader to arrive finds it
    // kDynamic = false keeps round 1's static order (CTA b: tiles b, b + grid, ...: a stage's next tile
    // follows from the one it held).  The K <= 8 shared-scan kernels are the one family that
    // measured FASTER with it (8-20 %, profiles/r02_ab_static_k8.txt); -DSSS_RING_STATIC=1 forces it
    // everywhere for A/B runs.
#ifndef SSS_RING_STATIC
#define SSS_RING_STATIC 0
#endif
    __device__ __forceinline__ uint64_t draw() { return atomicAdd(tickets, 1ull); }

    // one thread: start the copy of `tile` into `stage` (or record the end)
    __device__ __forceinline__ void fill(int stage, uint64_t tile)
    {
        if (tile < num_tiles) {  // (static order past the end: kNoTile + k * grid wraps to a small number only after
                                 //  2^64 tiles; tile_of is never kNoTile when a stage is released)
            tile_of[stage] = tile;
            const uint64_t off = tile * (uint64_t)kTileBytes;
            const uint64_t left = src_bytes - off;
            const uint32_t bytes = left < (uint64_t)kTileBytes ? (uint32_t)left : (uint32_t)kTileBytes;
            if (policy == 0) policy = l2_policy_evict_first();
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(base + stage * kStageBytes, src + off, bytes, &full[stage], policy);
        } else {
            tile_of[stage] = kNoTile;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
        }
    }

    // all threads; returns after the barriers are initialised and the first kStages copies are issued
    __device__ __forceinline__ void start()
    {
        if (threadIdx.x == 0) {
            for (int s = 0; s < kStages; s++) {
                mbar_init(&full[s], 1);
                released[s] = 0;
            }
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && (!kDynamic || SSS_RING_STATIC)) {
            for (int s = 0; s < kStages; s++) fill(s, blockIdx.x + (uint64_t)s * gridDim.x);
        } else if (threadIdx.x == 0) {
            uint64_t t[kStages + 2];
#pragma unroll
            for (int s = 0; s < kStages + 2; s++) t[s] = draw();  // independent atomics: one round trip for all
#pragma unroll
            for (int s = 0; s < kStages; s++) fill(s, t[s]);
            in_hand = t[kStages];      // release():      the tickets of the next two refills (even / odd iter)
            in_hand2 = t[kStages + 1];
            pend_even = clamp_ticket(t[kStages]);      // release_warp(): the same two, carried by this lane's first arrivals
            pend_odd = clamp_ticket(t[kStages + 1]);
        }
    }

    // all threads: wait for the stage of iteration `iter`; returns the tile it holds, kNoTile = no more work
    __device__ __forceinline__ uint64_t acquire(uint32_t iter)
    {
        const int stage = iter % kStages;
        mbar_wait(&full[stage], (iter / kStages) & 1u);
        return tile_of[stage];
    }

    // Barrier-free release for kernels that copy their part of the tile into registers first and
    // compute afterwards: every warp calls this once it has READ its words of the stage (pass one of
    // the loaded values as `dep` so the call cannot be scheduled above the loads).  The last of the
    // CTA's warps to arrive refills the stage itself; nobody waits for anybody.
    __device__ __forceinline__ void release_warp(uint32_t iter, uint32_t dep)
    {
        __syncwarp();
        if constexpr (!kDynamic || SSS_RING_STATIC) {  // static order: the stage's next tile follows from the one it held
            if ((threadIdx.x & 31) == 0) {
                const int stage = iter % kStages;
                uint32_t old;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;  // after %3"
                             : "=r"(old)
                             : "r"(smem_u32(&released[stage])), "r"(1u), "r"(dep)
                             : "memory");
                if (old == (blockDim.x >> 5) - 1) {
                    released[stage] = 0;
                    fence_proxy_async();
                    fill(stage, tile_of[stage] + (uint64_t)kStages * gridDim.x);
                }
            }
            return;
        }
        if ((threadIdx.x & 31) == 0) {
            // This lane's arrival carries, in the counter's high half, the ticket it drew at its last refill
            // of the same parity (two tiles ago or more) -- so the ticket reaches the next refiller through
            // the arrival atomic itself: no second shared-memory location, no fence.  (A first version
            // published it with a store + __threadfence_block: the fence waits for the lane's outstanding
            // global stores and cost the K <= 8 kernels 10 %.)
            uint32_t add = 1u;
            if (iter & 1u) {
                if (pend_odd != kNoTicket) add |= pend_odd << kCountBits;
                pend_odd = kNoTicket;
            } else {
                if (pend_even != kNoTicket) add |= pend_even << kCountBits;
                pend_even = kNoTicket;
            }
            const int stage = iter % kStages;
            uint32_t old;
            // `dep` is an unused operand: it only forces the loads that produced it to have completed
            asm volatile("atom.shared.add.u32 %0, [%1], %2;  // after %3"
                         : "=r"(old)
                         : "r"(smem_u32(&released[stage])), "r"(add), "r"(dep)
                         : "memory");
            old += add;
            if ((old & ((1u << kCountBits) - 1u)) == (blockDim.x >> 5)) {  // the last reader: refill with the ticket the arrivals carried
                released[stage] = 0;  // next touched only after the refilled tile has landed
                fence_proxy_async();
                const uint64_t tile = old >> kCountBits;
                fill(stage, tile);
                if (tile < num_tiles) {
                    if (iter & 1u) pend_odd = clamp_ticket(draw());
                    else pend_even = clamp_ticket(draw());
                } else {  // past the end: keep handing an invalid ticket on
                    if (iter & 1u) pend_odd = kTicketMask;
                    else pend_even = kTicketMask;
                }
            }
        }
    }

    // all threads, after the tile of iteration `iter` has been consumed
    __device__ __forceinline__ void release(uint32_t iter)
    {
        __syncthreads();
        if constexpr (!kDynamic || SSS_RING_STATIC) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                fill(iter % kStages, tile_of[iter % kStages] + (uint64_t)kStages * gridDim.x);
            }
            return;
        }
        if (threadIdx.x == 0) {
            fence_proxy_async();
            // two tickets in registers, one per parity of `iter`: the one drawn now is used two tiles later
            if (iter & 1u) {
                const uint64_t tile = in_hand2;
                fill(iter % kStages, tile);
                if (tile < num_tiles) in_hand2 = draw();
            } else {
                const uint64_t tile = in_hand;
                fill(iter % kStages, tile);
                if (tile < num_tiles) in_hand = draw();
            }
        }
    }

    // all threads, once, after the tile loop: the last CTA of the launch re-arms the ticket counter
    __device__ __forceinline__ void finish()
    {
        __syncthreads();  // no warp of this CTA draws a ticket after this
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&tickets[1], 1ull) == gridDim.x - 1) {
                tickets[0] = 0;
                tickets[1] = 0;
            }
        }
    }
};

// ---- code extraction --------------------------------------------------------------------------
//
// Code J of a 32-code block (W consecutive words in registers) moved to the TOP W bits of a 32-bit
// value.  Word index and shift distance are compile-time constants (sss::code_slot); the low 32-W
// bits hold leftovers of the preceding codes or zeros, which the aligned range compare ignores.
// Counterpart of the reference's pshufb byte gather + per-lane shift masks (commons :5-50).
template <int W, int J>
__device__ __forceinline__ uint32_t left_aligned_code(const uint32_t (&x)[W])
{
    constexpr CodeSlot s = code_slot(W, J);
    if constexpr (s.end_shift == 32) {
        return x[s.last_word];
    } else if constexpr (!s.straddles) {
        return x[s.last_word] << (32 - s.end_shift);
    } else {
        return __funnelshift_r(x[s.first_word], x[s.last_word], s.end_shift);
    }
}

template <int W, int J>
struct LeftAlignAll {
    static __device__ __forceinline__ void run(const uint32_t (&x)[W], uint32_t (&t)[32])
    {
        t[J] = left_aligned_code<W, J>(x);
        LeftAlignAll<W, J + 1>::run(x, t);
    }
};
template <int W>
struct LeftAlignAll<W, 32> {
    static __device__ __forceinline__ void run(const uint32_t (&)[W], uint32_t (&)[32]) {}
};

// ---- count reduction --------------------------------------------------------------------------

// Sum `v` over the CTA and add it to *global (one 64-bit atomic per CTA).  All threads call.
__device__ __forceinline__ void cta_add_u64(uint32_t v, unsigned long long* global)
{
    __shared__ uint32_t warp_sums[kWarps];
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < kWarps; i++) s += warp_sums[i];
        if (s) atomicAdd(global, s);
    }
}

// Count without a separate memset launch.  `slot` points at {accumulator, ticket}, both zero between
// launches (a context-owned pool, see capi.cu).  Every CTA adds its sum to the accumulator and takes
// a ticket; the CTA that draws the last ticket publishes the total to *result (stored, or added when
// `accumulate`), and re-zeroes the slot for its next user.
__device__ __forceinline__ void cta_finish_count(uint32_t v, unsigned long long* slot, unsigned long long* result,
                                                 bool accumulate)
{
    __shared__ uint32_t warp_sums[kWarps];
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < kWarps; i++) s += warp_sums[i];
        if (s) atomicAdd(&slot[0], s);
        __threadfence();
        const unsigned long long ticket = atomicAdd(&slot[1], 1ull);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const unsigned long long total = atomicExch(&slot[0], 0ull);
            slot[1] = 0;
            if (accumulate) atomicAdd(result, total); else *result = total;
        }
    }
}

}  // namespace sss
