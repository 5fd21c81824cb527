// tc.cuh - sm_100a building blocks written as inline PTX: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), UMMA shared-memory and instruction descriptors, and the
// host-side CUtensorMap encoder (driver entry point fetched at run time, so libgvit.so does not link libcuda).
//
// Shared-memory operand tiles all have one shape: [rows][64 bf16] = 128-byte rows, 128B-swizzled in
// 1024-byte (8-row) atoms - exactly what a TMA box of {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes.
// The same tile serves as a K-major operand (rows = M or N, the 64 columns = K) or as an MN-major
// operand (rows = K, the 64 columns = N); only the descriptor differs.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gvit {
namespace tc {

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// NOTE on single-thread roles (TMA producer, MMA issuer): enter them with
//     const int warp = __shfl_sync(~0u, threadIdx.x >> 5, 0);  if (warp == ROLE) { if (elect_one()) { ...loop... } }
// i.e. a warp-uniform role test followed by ONE elect.sync around the whole loop.  Measured on B200 (tools/mma_bench.cu):
// a tight single-thread issue loop costs N/2+10 cycles per tcgen05.mma with A in TMEM and N/2+41 with A in shared
// memory; an `if (lane == 0)` role made the compiler wrap every UTCHMMA in an ELECT/R2UR.BROADCAST/BRA.U.ANY waterfall
// (~125 cycles per MMA), and an elect per MMA still costs ~96.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// single arrival issued by one elected lane of a converged warp
__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) {
  if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the launch with an error, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
      printf("gvit: mbarrier wait timed out (block %d,%d,%d thread %d, barrier at shared offset %u, parity %u)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA ----
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 eviction-priority hints (the encodings createpolicy.fractional.L2::evict_{first,last}.b64 with fraction 1.0 returns;
// same constants as cute::TMA::CacheHintSm90).  evict_last keeps a tile that the SAME kernel re-reads a few microseconds
// later resident while other CTAs stream through L2; evict_first marks data nobody re-reads.
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void st_global_hint(void* p, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}
// same, delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) of every CTA of the
// cluster whose bit is set in cta_mask: ONE L2 read feeds several SMs
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// tile -> L2 only (no shared-memory destination, no barrier): warms L2 for a later tma_load of the same box
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// shared memory -> global tile store (rows outside the tensor are clipped by the TMA unit)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed stores have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recently committed group have (two staging buffers used in turn)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// the committed stores are complete (before the CTA exits)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05: TMEM allocation (one full warp executes these) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tcgen05.mma, kind::f16 (bf16 x bf16 -> fp32 in TMEM), issued by ONE thread ----
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// A operand read from TMEM (bf16 pairs packed per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cta_group::2 : one tcgen05.mma spans the two CTAs of a cluster pair (M = 256).  Each CTA stages ITS 128 rows of A and
// ITS half of B's N extent at the same shared-memory offsets; the leader (cluster rank 0) issues the instruction and the
// hardware reads both shared memories, so a CTA fills and feeds only HALF of B per tile.  Both CTAs allocate / free TMEM
// with the .cta_group::2 forms (same warp id, same slot address); barriers living in the leader are reached from the peer
// through a mapa-translated shared::cluster address. ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(rank));
  return r;
}
// remote (or local) arrival on an mbarrier given by its shared::cluster address.  Default semantics (.release at CTA scope):
// the arrivals here order tensor-memory reads (tcgen05.fence::before_thread_sync precedes them), not generic-proxy data
// another CTA will read - `.release.cluster` compiled to MEMBAR.ALL.CTA + ERRBAR per arrival, 13 % of the fused fc1
// kernel's stall samples (profiles/r4d_ncu_fc1_2sm.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// tile load into THIS CTA's shared memory whose completion bytes are counted on the mbarrier at `bar_cluster_addr`
// (the leader's full barrier of the stage)
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
__device__ __forceinline__ void umma_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// arrives (once all previously issued pair MMAs have completed) on the mbarrier at this shared-memory offset in every CTA
// of cta_mask
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- tcgen05.ld : warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); thread t gets lane base+t ----
// commit that arrives on the mbarrier at the same shared-memory offset in every CTA of the cluster selected by cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// ---- tcgen05.st : 16 consecutive 32-bit columns of this thread's lane (used to hand bf16x2-packed operands to a
// tcgen05.mma that reads A from tensor memory) ----
// two 32-column loads behind ONE wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t tb, float (&a)[32], float (&b)[32]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[32 + i]); }
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_lane_base(uint32_t taddr, int warp) { return taddr + (static_cast<uint32_t>((warp & 3) * 32) << 16); }

// ---- descriptors ----
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), majors at 15/16,
// N>>3 at 17-22, M>>4 at 24-28   [cute/arch/mma_sm100_desc.hpp, InstrDescriptor]
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// shared-memory matrix descriptor for a 128B-swizzled [rows][64 bf16] tile; same encoding for K-major
// (8-row atoms stacked along M/N every SBO bytes) and MN-major (8-row atoms stacked along K every SBO bytes;
// LBO would step between 64-wide MN atoms and is unused for a single 64-column tile).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  constexpr uint64_t SBO = 1024 >> 4, LBO = 1;
  return static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4) | (LBO << 16) | (SBO << 32) | (1ull << 46) | (2ull << 61);
}
// same, for an MN-major operand that spans several 64-element atoms along M/N: lbo_bytes = distance between them
__device__ __forceinline__ uint64_t make_sdesc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  constexpr uint64_t SBO = 1024 >> 4;
  return static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) | (SBO << 32) |
         (1ull << 46) | (2ull << 61);
}
// byte offset of element (row, col) inside a swizzled [rows][64 bf16] tile (col multiple of 8 -> 16-byte chunk)
__device__ __forceinline__ uint32_t swz128(int row, int col) {
  return static_cast<uint32_t>(row) * 128u + ((((static_cast<uint32_t>(col) >> 3) ^ (static_cast<uint32_t>(row) & 7u)) << 4));
}

// ---- work items of a persistent CTA pair (agg4_tc.cu, graph_bwd_pair_tc.cu) ----
// Full rounds: image cid + it * ncl, every output chunk.  The images left over after the full rounds
// (r = B mod ncl) would occupy r of the ncl pairs for a whole round; when 2 r <= ncl each of them is split between two pairs -
// both run the per-image phases (Z phase / Gram product, extraction, build), each takes one half of the 128-feature output
// chunks, and only the first half ("primary") writes the per-image side outputs (saved weights, saved Z, CLS row, dvals).
struct Item { int b, c0, c1; bool primary; };
__device__ __forceinline__ bool get_item(int it, int cid, int ncl, int B, int nchunk, Item& I) {
  const int R = B / ncl, r = B - R * ncl;
  I.c0 = 0; I.c1 = nchunk; I.primary = true;
  if (it < R) { I.b = cid + it * ncl; return true; }
  if (it > R || r == 0) return false;
  if (2 * r <= ncl && (nchunk & 1) == 0) {
    if (cid >= 2 * r) return false;
    I.b = R * ncl + (cid >> 1);
    I.primary = (cid & 1) == 0;
    I.c0 = (cid & 1) * (nchunk >> 1);
    I.c1 = I.c0 + (nchunk >> 1);
    return true;
  }
  if (cid >= r) return false;
  I.b = R * ncl + cid;
  return true;
}

// ---- neighbour indices from an adjacency list ----
// In range by contract (gvit_knn_fwd writes them), but the lists are caller-supplied pointers at the C ABI: an out-of-range
// value must never become a shared-memory scatter.  Product builds skip such an edge; -DGVIT_DEBUG_BOUNDS builds trap with a message.
__device__ __forceinline__ bool nb_ok(int i, int n) {
  const bool ok = static_cast<unsigned>(i) < static_cast<unsigned>(n);
#ifdef GVIT_DEBUG_BOUNDS
  if (!ok) {
    printf("gvit: neighbour index %d outside [0, %d) (block %d thread %d)\n", i, n, blockIdx.x, threadIdx.x);
    __trap();
  }
#endif
  return ok;
}

// ---- optional event trace (builds with -DGVIT_TRACE only; see tools/trace_kernel.py) ----
// CTA 0 records (warp, event id, clock64) tuples into a global buffer set by gvit_debug_set_trace(); used to see
// where the warp-specialised pipelines wait.  Compiled out of the product library.
#ifdef GVIT_TRACE
static __device__ unsigned long long* g_trace_buf = nullptr;
static __device__ unsigned int g_trace_cap = 0;      // entries per warp region (16 regions)
// no atomics: every traced warp owns a region of the buffer and a cursor in a register, so an event costs one
// CS2R + one fire-and-forget store
__device__ __forceinline__ void trace_event(int id, unsigned int& cursor) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && g_trace_buf != nullptr &&
      cursor < g_trace_cap) {
    const unsigned int warp = threadIdx.x >> 5;
    g_trace_buf[(warp & 15) * g_trace_cap + cursor] =
        (static_cast<unsigned long long>((warp << 8) | (id & 0xff)) << 44) |
        (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFFULL);
    ++cursor;
  }
}
// every CTA's begin (0) / end (1) in %globaltimer nanoseconds, behind the 16 event regions: launch skew, imbalance, tail
__device__ __forceinline__ void span_mark(int which) {
  if (threadIdx.x == 0 && g_trace_buf != nullptr && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace_buf[16 * g_trace_cap + 2 * blockIdx.x + which] = t;
  }
}
#define GVIT_SPAN(which) ::gvit::tc::span_mark(which)
#define GVIT_TRACE_DECL unsigned int gvit_trc = 0;
#define GVIT_TR(id) ::gvit::tc::trace_event(id, gvit_trc)
#define GVIT_TRACE_SETTER(NAME)                                                                   \
  extern "C" GVIT_API int NAME(void* buf, unsigned int cap) {                                      \
    unsigned long long* b = static_cast<unsigned long long*>(buf);                                 \
    if (cudaMemcpyToSymbol(::gvit::tc::g_trace_buf, &b, sizeof(b)) != cudaSuccess) return 4;       \
    if (cudaMemcpyToSymbol(::gvit::tc::g_trace_cap, &cap, sizeof(cap)) != cudaSuccess) return 4;   \
    return 0;                                                                                      \
  }
#else
#define GVIT_TR(id) ((void)0)
#define GVIT_SPAN(which) ((void)0)
#define GVIT_TRACE_DECL
#define GVIT_TRACE_SETTER(NAME)
#endif

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side: CUtensorMap for a bf16 tensor viewed as {inner, rows, batch} with box {64, box_rows, 1}
// ------------------------------------------------------------------------------------------------
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows);
// same for an fp32 tensor: box {32, box_rows, 1} (128-byte rows, 128-byte swizzle)
int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows);

}  // namespace gvit
