// sm_100a primitives used by the tensor-core kernels: mbarrier, bulk async copy
// (TMA engine, UBLKCP), TMEM allocation, tcgen05.mma / commit / ld, UMMA shared-memory and
// instruction descriptors.  Everything is inline PTX; no CUTLASS/CuTe types.
//
// Operand layout used throughout ("K-major, 128-byte swizzle"): a slab holds R rows of 64
// fp16 (= 128 bytes); row r lives at byte  r*128, and inside each aligned 8-row / 1024-byte
// group the 16-byte chunk c of row r is stored at chunk position  c ^ (r & 7).  Slabs are
// 1024-byte aligned so address bits [4,7) ^ [7,10) is exactly that permutation.  One slab
// covers 64 values of K; a tcgen05.mma consumes K=16 (32 bytes), so the four K-steps of a
// slab are addressed by advancing the descriptor start address by 32 bytes.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace kws {
namespace tc {

constexpr int TILE_M = 128;        // rows of an accumulator tile == TMEM lanes
constexpr int SLAB_K = 64;         // fp16 elements per swizzled row
constexpr int ROW_BYTES = 128;
constexpr int A_SLAB_BYTES = TILE_M * ROW_BYTES;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of (row, 16-byte chunk) inside a swizzled slab
__host__ __device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t chunk) {
  return row * ROW_BYTES + ((chunk ^ (row & 7u)) << 4);
}

// ---------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint expires.  Without a
// hint the default limit is a few cycles, and a waiting role re-issues try_wait + branch every ~20 cycles:
// a dozen waiting warps then burn issue slots and LSU bandwidth that the working warps need.
constexpr uint32_t MBAR_SUSPEND_NS = 100000u;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(a), "r"(parity), "r"(MBAR_SUSPEND_NS)
      : "memory");
}

__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {   // bar_addr = shared-space address
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar_addr), "r"(parity), "r"(MBAR_SUSPEND_NS)
      : "memory");
}

// Non-blocking probes of two mbarriers at once: the two round trips to the barrier unit (~150 cycles each,
// even when the phase completed long ago) overlap; bit 0 / bit 1 of the result = phase 1 / 2 complete.
__device__ __forceinline__ uint32_t mbar_test2(uint64_t* bar1, uint32_t parity1, uint64_t* bar2, uint32_t parity2) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p1, p2;\n\t"
      ".reg .b32 r1, r2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p1, [%1], %2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p2, [%3], %4;\n\t"
      "selp.u32 r1, 1, 0, p1;\n\t"
      "selp.u32 r2, 2, 0, p2;\n\t"
      "or.b32 %0, r1, r2;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar1)), "r"(parity1), "r"(smem_u32(bar2)), "r"(parity2)
      : "memory");
  return ok;
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------- bulk async copy global -> shared (TMA engine, no tensor map) ----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// same copy delivered to the same CTA-relative offsets (data and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// ---------------- thread-block clusters ----------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {               // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same CTA-relative address in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}

// ---------------- TMA tiled tensor load (2-D tensor map, no multicast) ----------------
// c0 = coordinate in the innermost (contiguous) dimension, c1 = row; out-of-range elements read as 0.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// TMA tiled tensor stores shared -> global (bulk async-group completion); elements outside the
// tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, int c0, int c1, int c2, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {      // <= N groups still reading their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// fire-and-forget prefetch of a contiguous global range into L2 (16-byte aligned address and size)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------- TMEM ----------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 2 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two fp32 -> packed fp16x2 with ReLU fused into the conversion (lo = first element in memory)
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---------------- UMMA descriptors ----------------
// shared-memory operand descriptor: K-major, SWIZZLE_128B, 8-row group stride (SBO) 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                               // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // SBO            [32,46)
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                               // SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: A,B fp16 (format 0) or bf16 (format 1), D fp32, K-major both
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n, int ab_format) {
  return (1u << 4) | (static_cast<uint32_t>(ab_format) << 7) | (static_cast<uint32_t>(ab_format) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Pinning kernel parameters in registers: parameters live in the constant bank, and ptxas prefers
// re-loading them (LDC / LDCU, tens of cycles each, serialised by the surrounding mbarrier waits) to
// keeping them live across a loop; in the single-warp pipeline roles those reloads were a large part
// of the per-slab latency.  The kernels add a zero that is read from shared memory at run time
// (`+ z`): the sum cannot be rematerialised from the constant bank, so it stays in a register.

// The MMA warps run their loops with all 32 lanes converged and elect one lane only around the
// tcgen05 instructions: addresses, stage counters and descriptors then live in uniform registers
// (a loop under `if (lane == 0)` is divergent code, and the compiler re-broadcasts every descriptor
// word through ELECT / R2UR in front of each MMA -- ~50 dependent instructions per tcgen05.mma,
// which made the ISSUE of the MMAs, not their execution, the bottleneck of the narrow layers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// low word of umma_desc_sw128 (start address + LBO); the high word is the constant below, and the
// next K step / slab / stage is an integer add on the low word
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
}
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI)
      : "memory");
}
// One K slab (4 x K=16) of MMAs plus up to two commits as ONE predicated instruction sequence: every
// lane computes the descriptors (uniform), the elected lane issues.  No divergent region, no per-MMA
// descriptor re-broadcast; bar1 / bar2 = shared-memory addresses of mbarriers to commit to (0 = none).
__device__ __forceinline__ void umma_slab4_commit(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                  uint32_t accumulate_first, uint32_t bar1, uint32_t bar2) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ax, bx;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.and.b32 p1, %5, 0, pe;\n\t"
      "setp.ne.and.b32 p2, %6, 0, pe;\n\t"
      "mov.b64 da, {%1, %7};\n\t"
      "mov.b64 db, {%2, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.u32 ax, %1, 2;\n\t"
      "add.u32 bx, %2, 2;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, %1, 4;\n\t"
      "add.u32 bx, %2, 4;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, %1, 6;\n\t"
      "add.u32 bx, %2, 6;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "@p1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@p2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(bar1), "r"(bar2), "r"(UMMA_DESC_HI)
      : "memory");
}
// Same for an A operand in the MN-major SW128 layout (rows = one k value x 64 contiguous M elements, 8 k per 1024-byte
// atom, LBO between the 64-element M blocks, SBO between the 8-k groups): the K = 16 steps of a slab advance the A start
// address by a_step (= 2 SBO, in 16-byte units) and the A descriptor has its own high word.
__device__ __forceinline__ uint32_t umma_desc_lo_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__host__ __device__ constexpr uint32_t umma_desc_hi_mn(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_slab4_commit_mn(uint32_t tmem_d, uint32_t a_lo, uint32_t a_step, uint32_t a_hi, uint32_t b_lo,
                                                     uint32_t idesc, uint32_t accumulate_first, uint32_t bar1, uint32_t bar2) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ax, bx;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.and.b32 p1, %5, 0, pe;\n\t"
      "setp.ne.and.b32 p2, %6, 0, pe;\n\t"
      "mov.b64 da, {%1, %9};\n\t"
      "mov.b64 db, {%2, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.u32 ax, %1, %8;\n\t"
      "add.u32 bx, %2, 2;\n\t"
      "mov.b64 da, {ax, %9};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, ax, %8;\n\t"
      "add.u32 bx, %2, 4;\n\t"
      "mov.b64 da, {ax, %9};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, ax, %8;\n\t"
      "add.u32 bx, %2, 6;\n\t"
      "mov.b64 da, {ax, %9};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "@p1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@p2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(bar1), "r"(bar2), "r"(UMMA_DESC_HI), "r"(a_step), "r"(a_hi)
      : "memory");
}
// same with a run-time number of K steps (1..4): the last K slab of a contraction whose K is not a multiple of 64
__device__ __forceinline__ void umma_slab_commit(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                 uint32_t accumulate_first, uint32_t ksteps, uint32_t bar1, uint32_t bar2) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt, p1, p2, q1, q2, q3;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ax, bx;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.and.b32 p1, %5, 0, pe;\n\t"
      "setp.ne.and.b32 p2, %6, 0, pe;\n\t"
      "setp.gt.and.u32 q1, %8, 1, pe;\n\t"
      "setp.gt.and.u32 q2, %8, 2, pe;\n\t"
      "setp.gt.and.u32 q3, %8, 3, pe;\n\t"
      "mov.b64 da, {%1, %7};\n\t"
      "mov.b64 db, {%2, %7};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.u32 ax, %1, 2;\n\t"
      "add.u32 bx, %2, 2;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, %1, 4;\n\t"
      "add.u32 bx, %2, 4;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ax, %1, 6;\n\t"
      "add.u32 bx, %2, 6;\n\t"
      "mov.b64 da, {ax, %7};\n\t"
      "mov.b64 db, {bx, %7};\n\t"
      "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "@p1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@p2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(bar1), "r"(bar2), "r"(UMMA_DESC_HI), "r"(ksteps)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {      // all lanes call; the elected lane commits
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar)
      : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

}  // namespace tc
}  // namespace kws
