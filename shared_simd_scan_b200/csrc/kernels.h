// kernels.h -- internal launch interface between the C-ABI shim (capi.cu) and the kernel families.
// Every launcher dispatches the runtime width to a template<int W> kernel through a 32-entry table
// (the reference fixes the width at compile time: `#define BITS_NEEDED 9`, src/simd_scan.hpp:12).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/simd_scan_commons.hpp"

namespace sss {

constexpr uint32_t kTicketSlots = 1024;  // launches in flight that may draw tiles concurrently

struct Device {
    int ordinal;
    int sm_count;
    int smem_optin;  // max dynamic shared memory per CTA
    // tile-ticket counters of the persistent kernels (common.cuh TileRing): kTicketSlots x {next tile,
    // CTAs finished}, zero between launches (every launch re-arms its slot itself)
    unsigned long long* ticket_pool = nullptr;
    mutable uint32_t ticket_cursor = 0;
    // scan.cu: columns of at most this many tiles per resident CTA take the ring-less kernel (SSS_SCAN_DIRECT_TILES
    // in the environment at sss_ctx_create overrides it; 0 = always the persistent ring kernel -- the parity
    // tests run both)
    uint32_t scan_direct_tiles = 8;
    // scan.cu: launch the ring-less scan kernel with programmatic stream serialization (SSS_PDL=0 in the
    // environment at sss_ctx_create switches it off): its CTAs are scheduled while the previous kernel drains and
    // wait in griddepcontrol.wait, at the top of the kernel, until it has completed -- stream order as before,
    // without the launch gap between back-to-back scans
    bool pdl = true;
    unsigned long long* next_tickets() const { return ticket_pool + 2 * (size_t)(ticket_cursor++ % kTicketSlots); }
};

// Where a kernel's match count goes: `result` (device, may be null = no count) is stored or added to by
// the last CTA; `slot` = {accumulator, ticket}, zero between launches (context-owned pool), so no
// memset launch is needed per scan.
struct CountTarget {
    unsigned long long* result;
    unsigned long long* slot;
    bool accumulate;
    // scans only: when set (zeroed by the caller), the scan also adds every 1024-word segment's matches to
    // seg_count[segment] -- the first pass of the position list (positions.cu) without reading the bitvector
    uint32_t* seg_count = nullptr;
};
constexpr int kPosSegShift = 10;  // a position-list segment = 2^10 result words = 32768 codes (positions.cu)

// scan.cu -- fused unpack + range compare -> bitvector (+count)
cudaError_t launch_scan_range(const Device& dev, const void* packed, uint64_t n, int w, AlignedRange r,
                              uint32_t* bitvec, const CountTarget& count, cudaStream_t stream);

// the same scan with lo/hi in device memory (one-predicate shared scans); an empty range writes zeros
cudaError_t launch_scan_range_dev(const Device& dev, const void* packed, uint64_t n, int w, const uint32_t* d_lo,
                                  const uint32_t* d_hi, uint32_t* bitvec, const CountTarget& count, cudaStream_t stream);

// unpack.cu -- packed -> int32 per code
cudaError_t launch_unpack(const Device& dev, const void* packed, uint64_t n, int w, int32_t* out,
                          cudaStream_t stream);

// predindex.cu -- the K bounds compiled into a lookup structure (see predindex.cuh)
size_t pred_index_bytes();
cudaError_t launch_build_pred_index(const uint32_t* d_lo, const uint32_t* d_hi, int K, int w, void* d_index,
                                    cudaStream_t stream);
cudaError_t launch_counts_from_histogram(void* d_index, int K, unsigned long long* d_counts, cudaStream_t stream);

// shared.cu -- K predicates in one pass -> K bitvectors.  Match counts: for K <= kDirectCountMaxK the
// kernel ADDS them to counts[0..K) itself; above they go to the index's interval histogram and
// launch_counts_from_histogram finishes the job.
constexpr int kDirectCountMaxK = 8;
cudaError_t launch_shared_scan(const Device& dev, const void* packed, uint64_t n, int w, const void* index, int K,
                               uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                               cudaStream_t stream);

// Bitvector layout, device-resident entry point: the largest number of predicates to hand to one
// launch_shared_scan call so that the kernel's row table stays in shared memory (K itself when it
// fits or nothing smaller would); the caller loops over the groups.
int shared_scan_predicate_group(int w, int K);

// narrow.cu -- K <= kNarrowMaxK predicates over codes of at most kNarrowMaxW bits, either output layout:
// a table indexed by a GROUP of adjacent codes replaces the per-code interval lookup.  Adds the match
// counts to counts[0..K) itself (may be null).
constexpr int kNarrowMaxK = 8;
constexpr int kNarrowMaxW = 13;  // == the direct-locator widths (predindex.cuh kDirectMaxW)
constexpr int kFewPredicatesMaxW = 28;  // ... and up to here through the mask-carrying bucket table (narrow.cu)
cudaError_t launch_shared_scan_narrow(const Device& dev, const void* packed, uint64_t n, int w, const void* index,
                                      int K, void* out, uint64_t stride_words, bool linear,
                                      unsigned long long* counts, cudaStream_t stream);

// linear.cu -- K predicates in one pass -> [4*ceil(n/32)][K] byte matrix
cudaError_t launch_shared_scan_linear(const Device& dev, const void* packed, uint64_t n, int w, const void* index,
                                      int K, uint8_t* out, unsigned long long* counts, bool count_matches,
                                      cudaStream_t stream);

// pack.cu -- compressor and synthetic-column generator
cudaError_t launch_pack(const Device& dev, const uint32_t* codes, uint64_t n, int w, uint32_t* packed,
                        uint64_t packed_words, cudaStream_t stream);
cudaError_t launch_generate_packed(const Device& dev, uint32_t* packed, uint64_t packed_words, uint64_t n, int w,
                                   uint64_t seed, uint64_t first_index, cudaStream_t stream);

// consumers.cu -- what a column store does next with a scan result
cudaError_t launch_bitvec_combine(const Device& dev, int op, const uint32_t* a, const uint32_t* b, uint64_t n,
                                  uint32_t* out, const CountTarget& count, cudaStream_t stream);
cudaError_t launch_unpack_at(const Device& dev, const void* packed, int w, const unsigned long long* positions,
                             uint64_t num, int32_t* out, cudaStream_t stream);

// calib.cu -- HBM bandwidth for the access mixes of this path (read / write / copy / read:write = ratio:1)
cudaError_t launch_calibrate(const Device& dev, int mode, const void* a, void* b, uint64_t bytes, uint32_t ratio,
                             unsigned long long* sink, cudaStream_t stream);

// positions.cu -- bitvector -> ascending position list in three launches that never wait (count per 4 KB
// segment, group offsets, ordered expansion).  The workspace holds the group offsets and the segment counts;
// `epoch` is kept for the C ABI's bookkeeping (capi.cu counts it per context; the kernels carry no state
// between launches).
size_t positions_workspace_bytes(uint64_t n);
uint32_t positions_epoch_limit();
cudaError_t launch_positions(const Device& dev, const uint32_t* bitvec, uint64_t n, uint64_t base,
                             unsigned long long* positions, uint64_t capacity, unsigned long long* num,
                             void* workspace, uint32_t epoch, cudaStream_t stream, bool counted = false);
// the workspace's per-segment match counts: a scan launched with CountTarget::seg_count pointing here (zeroed
// first) makes launch_positions(..., counted = true) skip its own count pass over the bitvector
uint32_t* positions_seg_counts(void* workspace, uint64_t n);
size_t positions_seg_count_bytes(uint64_t n);

}  // namespace sss
