// Tensor-core (tcgen05 / TMEM) jet-MLP kernel for sm_100a, TF32 operands with FP32 accumulation.
//
// Same contract as the FP32 kernel (jet_fp32.cu): per tile of collocation points, forward value +
// tangent jets through every Linear+tanh layer, fused PDE-residual epilogue, reverse sweep to the
// flat weight gradient.  Here the 256x256 hidden layers -- >99 % of the FLOPs -- run on the 5th-gen
// tensor cores:
//
//   CTA pair        two CTAs (one cluster) work on two tiles in lock step; the leader's MMA warp issues
//                   tcgen05.mma.cta_group::2 for both (M = 256 = 2 x 128 rows); every B operand is split in
//                   halves, each CTA streaming only its own
//   tile            32 points x 4 jet rows = 128 rows per CTA
//   forward  l      D[256 x 256] = A[256 x 256] * W_l^T          A: smem (K-major), B: W half-images via TMA
//   adjoint  l      D[256 x 256] = Zbar[256 x 256] * W_l         A: smem (K-major), B: W^T half-images via TMA
//   weight grad l   D[256 x 256] = A_in^T * Zbar (the TRANSPOSED gradient), K = the 256 rows of BOTH tiles; CTA r ends up with
//                   dW columns [128 r, 128 r + 128) as TMEM lanes   A, B: MN-major (SWIZZLE_128B_BASE32B) spill pieces via TMA
//   last layer      forward D[256 x 32] = A * W_last^T (N = 32, zero-padded); reverse Abar = seeds * W_last as
//                   ONE K = 8 MMA whose A operand is the K-major output / seed image
//   accumulators live in TMEM (512 columns), read back with tcgen05.ld for the tanh / jet / adjoint
//   epilogues; only the d->256 layer stays on the FP32 pipes.  Forward jobs alternate between TMEM columns
//   0..255 and 256..511 and consume the operand image in four 64-feature slices, so layer l+1 starts while the
//   epilogue of layer l is still running; in the reverse sweep columns 256..511 hold the weight-gradient
//   accumulator.
//
// Row order inside a tile: row m = 32*sp + 8*j + pp holds jet j of point 8*sp + pp (sp = TMEM
// subpartition).  A tcgen05.ld.16x256b hands thread T of a warp the rows pp = T/4 and pp + 8 and two
// adjacent columns, so two loads (lanes 0-15, 16-31) give ONE thread all four jets of a point for its
// features: the tanh' coupling between the value and the tangent rows is thread-local (no shuffles) and
// tanh is evaluated once per (point, feature).
//
// Warp roles (640 threads, one CTA per SM, persistent over tile pairs):
//   warps 0-15 workers: layer 0, epilogues (TMEM -> registers -> operand image in smem / spill), drains (RED); 104 registers
//   warp  16   producer: one lane per ring stage issues the TMA copies of weight half-images and spill pieces
//   warp  17   leader: one thread issues every tcgen05.mma / tcgen05.commit of the pair; owns the TMEM allocation (the
//              follower's warp 17 only relays "my stage has landed" in the -DNO_TMAP_RING build)
//   warps 18-19 spare (setmaxnreg moves registers between whole warpgroups: the control warpgroup 16-19 runs on 64)
//
// Operand image in shared memory (K-major, no swizzle): element (row m, feature f) at byte
//   (f/4)*OP_LBO + m*16 + (f%4)*4, OP_LBO = 128*16 + 32: the canonical K-major UMMA layout with
//   LBO = OP_LBO, SBO = 128; the 32-byte pad makes the 16-byte jets-in-thread accesses conflict-free.
// Spill images in global memory (128 KB each, per CTA: a_0 .. a_{L-3}, then two Zbar buffers used alternately): the two
//   operands of the weight-gradient job of layer l are Zbar_l (written in the reverse sweep) and a_{l-1} (written in the
//   forward sweep).  An image = [16-row chunk (8)][feature half (2)][8 KB piece]; a piece = [4 panels of 32 features][16 rows]
//   [128 B] with the 32-byte units of a row XOR-swizzled by (row % 4): exactly the MN-major SWIZZLE_128B_BASE32B shared-memory
//   layout (LBO 2048 between panels, SBO 512 between 4-row atoms), the only MN-major layout kind::tf32 accepts
//   (tools/umma_mn_probe.cu).  It is row-major per panel, so the epilogue threads write it straight from registers in full
//   32-byte sectors and read a_{l-1} back the same way in the reverse sweep; a ring stage of the weight-gradient job = the
//   Zbar piece and the a piece `rank` of one 16-row chunk (two 8 KB copies on one barrier) -- no transposition pass anywhere.
//
// Split-operand mode (X3 = true, PINN_PREC_TF32X3): FP32-grade results on the same tensor pipe.  Every operand x is
// carried as x = hi + lo with hi = tf32(x), lo = tf32(x - hi) (22 mantissa bits together).  A tile holds 16 points;
// the hi and lo parts of a jet are SEPARATE ROWS of the same 128-row operand image:
//   row m = 32*sp + 8*a + pp,   a = 2*jj + h (h = 0 hi, 1 lo),   pp = 4*jh + q,   jet j = 2*jh + jj,   point 4*sp + q
// so the MMA shape, the descriptors, the ring and the spill layout are unchanged and
//   forward / adjoint l   D = [A_hi; A_lo] * B_hi  then  += [A_hi; A_lo] * B_lo  (weights streamed as hi and lo images);
//                         the epilogue adds the hi and the lo row of a jet in FP32 -- the small terms accumulate
//                         in their own TMEM row -- i.e. (A_hi + A_lo)(B_hi + B_lo), all four products
//   weight grad l         the two K atoms of a 16-row spill piece are the hi rows (kk = 0) and the lo rows (kk = 1) of the
//                         same (point, jet) set: Zbar_hi^T A_hi + Zbar_hi^T A_lo + Zbar_lo^T A_hi by pairing the atoms through
//                         the descriptors -- no second spill image
// A thread's four row slots (a = 0..3) hold two jets (hi, lo) of ONE point; the partner lane (lane ^ 16) holds the other
// two, so the tanh' coupling costs a few warp shuffles.  tanhf instead of tanh.approx.
#include <stdlib.h>
// The ring is fed by 2-D tiled TMA copies (tensor maps over the packed weights and the spill slab, rows of 1 KB) with
// cta_group::2 completion: both CTAs' copies of a stage complete their bytes on the LEADER's barrier, so the issuer learns
// that the follower's half has landed without a relay lane and a remote arrive (-1.4 % / -2.3 % of the evaluation time).
// -DNO_TMAP_RING builds the 1-D bulk-copy ring with the relay instead.
#ifndef NO_TMAP_RING
#define TC_TMAP_RING
#endif
#ifdef TC_TMAP_RING
#include <cuda.h>
#ifndef TC_TMAP_L2PROMO
#define TC_TMAP_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
#endif

#include "common.cuh"
#include "residual.cuh"

namespace pinn {

constexpr int TC_H = 256;                 // hidden width handled by this kernel
constexpr int TC_M = 128;                  // rows per tile
constexpr int TC_TP = 32;                  // points per tile (TF32 mode; the split-operand mode carries 16, see below)
constexpr int TC_TP_X3 = 16;
constexpr int TC_WPS = 4;                  // worker warps per TMEM subpartition
constexpr int TC_WORKERS = 128 * TC_WPS;
constexpr int TC_THREADS = TC_WORKERS + 128;   // 16 worker warps + one control warpgroup: producer, issuer / relay, two spare warps
// Registers: with 20 warps every SM sub-partition holds five, so the launch bound is 96 per thread (16,384 / (5 x 32) = 102).
// setmaxnreg moves registers inside the CTA's own allocation (640 x 96): the control warpgroup shrinks to 64 per thread
// and hands 4 x 32 x 32 registers to the 16 worker warps, which grow to 104 (16 x 32 x 8): no spills in the epilogues and
// room to fetch the next layer's stored activations before the drain.
constexpr int TC_WORKER_REGS = 104;
constexpr int TC_CONTROL_REGS = 64;
constexpr int TC_WCOLS = TC_H / TC_WPS;    // columns of a 128x256 accumulator owned by one worker warp
constexpr int TC_NBLK = TC_WCOLS / 16;     // 16-column blocks per worker warp
constexpr int TC_PARTS = TC_WORKERS / TC_H;  // worker threads per feature in the thread-per-feature phases
constexpr int TC_STAGES = 5;                // ring: 5 x 16 KB (8 KB stages cost more per-stage barrier traffic than they gain)
constexpr int TC_STAGE_BYTES = 16384;
constexpr int TC_STAGE_FLOATS = TC_STAGE_BYTES / 4;
constexpr int TC_CHUNK16 = 4096;            // floats per 16 KB unit of the packed weight / edge images (= two ring stages)
constexpr int TC_PIECE = 2048;              // floats per spill piece: 16 rows x 128 features, MN-major swizzled (half a ring stage)
constexpr int OP_LBO = TC_M * 16 + 32;     // 2080
constexpr int OP_BYTES = (TC_H / 4) * OP_LBO;
constexpr int TC_IMG = TC_M * TC_H;        // floats per spill image (one quantity of one layer of one tile)
constexpr int TC_EDGE_W0 = 0;               // float offsets inside the edge block that follows the hidden-layer weight images
constexpr int TC_EDGE_WL = TC_H * 8;        //   W0 padded [H][8] | Wlast padded [8][H] | forward-last B images [rank 2][hi,lo][4096] |
constexpr int TC_EDGE_E1 = 2 * TC_H * 8;    //   reverse-last B images [rank 2][hi,lo][1024]   (the lo parts are used by the split-operand mode only)
constexpr int TC_EDGE_E2 = TC_EDGE_E1 + 4 * TC_CHUNK16;
constexpr int TC_EDGE_FLOATS = TC_EDGE_E2 + 4 * 1024;
constexpr int TC_WCHUNKS = TC_H * (TC_H / 2) * 4 / TC_STAGE_BYTES;   // 8 chunks per half-width weight image
constexpr int TC_MAX_HH = 7;               // hidden->hidden layers whose bias gradients are staged in shared memory (deeper ones: global atomics)

struct TcArgs {
  const float* params;
  const float* packed;   // per hidden->hidden layer 4 half-width weight images; then W0 padded [H][8], Wlast padded [8][H]
  const float* inputs;
  const float* targets;
  const float* mask_count;
  float* grad;
  double* sums;
  float* out;
  float* dout[PINN_MAX_DIRS];
  float* slab;            // per CTA: L-2 activation images a_0 .. a_{L-3}, then two alternating Zbar buffers
  long long slab_stride;  // floats per CTA
  long long n_points;
  int n_tiles;
  float inv_n_res;
  float inv_n_fid;
  float comp_dw;          // split-operand mode: accumulator-truncation compensation of the weight-gradient jobs (see x3_comp)
};

// The tensor core updates its FP32 accumulator with round-toward-zero, once per MMA instruction
// (tools/umma_acc_probe.cu: the TMEM result equals acc <- RZ(acc + exact sum of the 8 products) bit for bit; a K = 256
// contraction of positive numbers comes out 1.4e-6 low).  Truncation always pulls towards zero, so over the n instructions
// that accumulate into one TMEM word the expected loss is a fraction of an ulp of every partial sum -- to first order
// proportional to the partial sums themselves, hence to the result: about kappa * (n + 1) / 2 of it.  At TF32 noise
// (5e-4) that is invisible; at the 3xTF32 level it is the leading error (every layer shrinks its output by ~1.4e-6, a
// 256x8 net loses 1.6e-5 of its gradient norm with an angular error of only 2.6e-6).  The split-operand mode therefore
// scales what it feeds the accumulators by 1 + kappa (n + 1) / 2: folded into the packed weight images for the forward /
// adjoint jobs (n = 64 instructions per accumulator) and applied in the drain for the weight-gradient jobs (n = 48).
// kappa = the mean truncation loss per instruction, half an ulp, relative to the value: an ulp is 2^-23 / m of a number
// with mantissa m in [1,2), and E[1/m] = 1/(2 ln 2) for log-uniform m, so kappa = 2^-24 / (2 ln 2) = 4.30e-8.  The sweep in
// profiles/r2_x3_calib.log (tools/x3_calib.py, PINN_X3_KAPPA overrides the constant) puts the zero crossing of the scale
// error at 3.8e-8 .. 4.6e-8 for five nets of depth 2..8: with it the loss error drops from 1.7e-5 to <= 3e-6 and the
// gradient error from 1.6e-5 to <= 3e-6 (what is left is the angular part).
constexpr float TC_X3_KAPPA = 4.30e-8f;
inline float x3_kappa() {
  const char* e = getenv("PINN_X3_KAPPA");
  return e ? (float)atof(e) : TC_X3_KAPPA;
}
inline float x3_comp(int n_instr) { return 1.0f + x3_kappa() * 0.5f * (float)(n_instr + 1); }

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_WORKERS) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors as (low word, high word): only the low word (start address, LBO) varies per MMA, so the issue loop does
// 32-bit arithmetic and keeps two high words for all jobs -- the issuer runs in a 64-register warpgroup
__device__ __forceinline__ void umma_tf32_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair when all previously issued MMAs
// of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// no ordering of this thread's own earlier accesses: for pure hand-offs (the relay), where a releasing arrive would
// make every arrive wait for the round trip of the previous one
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 TMEM lanes x 32 columns; thread T gets, for u = 0..3: v[4u], v[4u+1] = (lane T/4,     cols 8u + 2(T%4) + {0,1})
//                                                         v[4u+2], v[4u+3] = (lane T/4 + 8, same cols)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float round_tf32(float x) {
  // cvt.rna.tf32.f32 (round to nearest, ties away from zero) on the sign-magnitude bit pattern: two ALU ops instead of
  // one XU op -- the epilogues are XU-bound (tanh + 16 conversions per 16 accumulators)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

#ifdef TC_TMAP_RING
// 2-D tiled TMA copy whose transaction bytes complete on an mbarrier of EITHER CTA of the pair (cta_group::2): the
// follower's copies signal the leader's "stage full" barrier directly -- no relay lane, no remote arrive
__device__ __forceinline__ void tma_load_2d_pair(void* dst_smem, const CUtensorMap* tm, int row, uint32_t mbar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst_smem)),
      "l"(tm), "r"(0), "r"(row), "r"(mbar_cluster)
      : "memory");
}
#define TC_TMAP_PARAMS , const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_rl, const __grid_constant__ CUtensorMap tm_p
#else
#define TC_TMAP_PARAMS
#endif
// UMMA shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" core matrices of 8 x 16 bytes)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;
}
// instruction descriptor: TF32 x TF32 -> F32, M x N, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float warp_sum_tc(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16 TMEM lanes x 16 columns, no wait: thread T gets, for u = 0..1:
//   r[4u], r[4u+1] = (lane T/4, cols 8u + 2(T%4) + {0,1}),   r[4u+2], r[4u+3] = (lane T/4 + 8, same cols)
__device__ __forceinline__ void tmem_ld_16x256b_x2_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float (&v)[8]) {   // thread = TMEM lane, 8 columns
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// MN-major operand descriptor, SWIZZLE_128B_BASE32B (layout type 1): LBO = stride between 32-element
// panels along M/N, SBO = stride between 4-row atoms along K
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return umma_desc(smem_addr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void st_global_v4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// L2 eviction-priority hints for the spill streams
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_global_v4_hint(float* p, float a, float b, float c, float d, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_global_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Output jets / adjoint seeds of a tile live in a K-major UMMA operand image [k/4 (2)][row (128)][4] (the A operand of
// the reverse last-layer MMA): element (row m, output column c) at float (c/4)*512 + m*4 + c%4.
__device__ __forceinline__ int outs_idx(int m, int c) { return (c >> 2) * 512 + m * 4 + (c & 3); }
// row of jet j of point p of a tile (split-operand mode: its hi row; the lo row is 8 further)
template <bool X3>
__device__ __forceinline__ int tile_row(int p, int j) {
  return X3 ? 32 * (p >> 2) + 16 * (j & 1) + 4 * (j >> 1) + (p & 3) : 32 * (p >> 3) + 8 * j + (p & 7);
}
template <bool X3>
struct TileJets {  // output jets of one point
  float* outs;
  int p;
  __device__ __forceinline__ int row(int j) const { return tile_row<X3>(p, j); }
  __device__ __forceinline__ float get(int col, int j) const { return outs[outs_idx(row(j), col)]; }
  __device__ __forceinline__ void set(int col, int j, float v) { outs[outs_idx(row(j), col)] = v; }
  __device__ __forceinline__ void add(int col, int j, float v) { outs[outs_idx(row(j), col)] += v; }
};

#if defined(PINN_TC_DEBUG) && !defined(PINN_TC_PHASES)
#define PINN_TC_PHASES
#endif
#ifdef PINN_TC_PHASES
#define TCT_DECL long long tct[16] = {0}; long long tct0 = clock64();
#define TCT(i) { const long long t_ = clock64(); tct[i] += t_ - tct0; tct0 = t_; }
#else
#define TCT_DECL
#define TCT(i)
#endif

#ifdef KO_RING
#define RINGWAIT(x)
#else
#define RINGWAIT(x) x
#endif
#ifdef PINN_TC_TRACE
// event trace of ONE tile pair of cluster 0 (leader CTA): (tag, clock64) pairs, read back with pinn_debug_trace()
__device__ long long tc_trace_buf[2 * 8192];
__device__ int tc_trace_n;
#define TR(tag) { if (tr_on) { tc_trace_buf[2 * (tr_base + tr_n)] = (tag); tc_trace_buf[2 * (tr_base + tr_n) + 1] = clock64(); ++tr_n; } }
#ifdef PINN_TC_TRACE_STAGES
#define TRS(tag) TR(tag)
#else
#define TRS(tag)
#endif
#define TR_SET(cond) tr_on = blockIdx.x == 0 && it == PINN_TC_TRACE && (cond); tr_n = 0;
#else
#define TR(tag)
#define TRS(tag)
#define TR_SET(cond)
#endif
template <bool BWD, bool X3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    jet_tc_kernel(const __grid_constant__ pinn_desc_t D, const __grid_constant__ TcArgs A TC_TMAP_PARAMS) {
  constexpr int TP = X3 ? TC_TP_X3 : TC_TP;       // points per tile
  constexpr int XP = X3 ? 2 : 1;                  // weight images per product (hi, lo)
  constexpr size_t LSTRIDE = (size_t)(2 * XP) * TC_H * TC_H;   // packed floats per hidden layer: [fwd hi][adj hi]([fwd lo][adj lo])
  constexpr size_t LO_OFF = (size_t)2 * TC_H * TC_H;           // offset of the lo images inside a layer's block
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* op = smem_raw;                                   // operand image (A or Zbar)
  unsigned char* ring = op + OP_BYTES;
  float* outs = reinterpret_cast<float*>(ring + TC_STAGES * TC_STAGE_BYTES);  // [128][8] output jets / seeds
  float* xin = outs + TC_M * 8;                                   // [32][8]
  float* db_s = xin + TC_TP * 8;                                  // [TC_MAX_HH][H] hidden-layer bias gradients
  double* red = reinterpret_cast<double*>(db_s + TC_MAX_HH * TC_H);  // [16]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + PINN_NSUMS);  // [S]
  uint64_t* empty = full + TC_STAGES;                             // [S]
  uint64_t* full_peer = empty + TC_STAGES;                        // [S] leader only: the follower's stage has landed
  uint64_t* op_ready = full_peer + TC_STAGES;
  uint64_t* edge_ready = op_ready + 4;   // op_ready[4]: one barrier per 64-feature slice (a parity-waited barrier must not
                                         // advance twice unseen, and the workers finish all four slices in a burst);
                                         // edge_ready: the adjoint seeds are in the output image (not sliced)
  uint64_t* mma_done = edge_ready + 1;
  uint64_t* slab_ready = mma_done + 1;
  uint64_t* zt_ready = slab_ready + 1;   // [2], alternating per Zbar spill: the workers may run one spill ahead of the
                                         // producer's wait, and a single parity-waited barrier must never advance twice unseen
  uint64_t* mma_done_b = zt_ready + 2;   // weight-gradient accumulator (TMEM columns 256..511) complete
  uint64_t* rb_free = mma_done_b + 1;    // ... and drained by the workers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rb_free + 1);
  float* w0s = reinterpret_cast<float*>(tmem_ptr + 4);            // [H][4]: columns 0..3 of W0, resident for the whole kernel

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef PINN_TC_TRACE
  bool tr_on = false;
  int tr_n = 0;
  const int tr_base = warp < TC_WORKERS / 32 ? 0 : (warp == TC_WORKERS / 32 ? 6144 : 512);   // workers | issuer | producer
#endif
  const int L = D.n_linear;
  const int d = D.widths[0], o = D.widths[L];
  const int NHH = L - 2;                      // hidden->hidden layers (tensor-core jobs per direction)
  const int kind = D.residual_kind;
  // CTA pair: the two CTAs of a cluster work on two tiles in lock step.  Forward / adjoint jobs are ONE
  // cta_group::2 MMA stream (M = 256: each CTA's 128 rows; each CTA streams only its 128-column half of the
  // weights); the weight-gradient job of a layer contracts over the rows of BOTH tiles, CTA r owning the Zbar
  // features [128 r, 128 r + 128) -- one drain per layer per CTA.
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  const int tile_pairs = (A.n_tiles + 1) >> 1;
  const int my_tiles = (tile_pairs - pair + n_pairs - 1) / n_pairs;   // tile pairs this cluster processes
  float* slab = A.slab + (long long)blockIdx.x * A.slab_stride;  // G_1 .. G_{L-2} of this CTA's tile
  const float* slab_pair[2] = {A.slab + (long long)(2 * pair) * A.slab_stride, A.slab + (long long)(2 * pair + 1) * A.slab_stride};
  const long long P0 = (long long)d * TC_H + TC_H;               // params of layer 0
  const long long PH = (long long)TC_H * TC_H + TC_H;            // params of a hidden->hidden layer
  const long long poffL = P0 + (long long)NHH * PH;              // params offset of the last layer
  const float* edge = A.packed + (size_t)NHH * LSTRIDE;          // edge-layer block of the packed weights
  const float* w0p = edge + TC_EDGE_W0;                          // [H][8]  W0[f][c], zero-padded (L1-resident)

  // ---------------- one-time setup ----------------
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
#ifdef SPLIT_FULL
      mbar_init(&full[s], 1);
#elif defined(TC_TMAP_RING)
      mbar_init(&full[s], 1);   // leader: its producer's expect_tx for the bytes of BOTH CTAs; the follower's barrier is unused
#else
      // leader: a stage is full when its own copy has landed (the producer's expect_tx arrive + the bytes) AND the follower
      // has relayed that its half has landed -- one barrier, one wait per stage in the issue loop
      mbar_init(&full[s], rank == 0 ? 2 : 1);
#endif
      mbar_init(&empty[s], 1);
      mbar_init(&full_peer[s], 1);
    }
    // cross-CTA barriers count one elected arrival per worker warp of each CTA
    for (int q = 0; q < 4; ++q) mbar_init(&op_ready[q], 2 * TC_WORKERS / 32);
    mbar_init(edge_ready, 2 * TC_WORKERS / 32);
    mbar_init(mma_done, 1);
    mbar_init(slab_ready, 2 * TC_WORKERS / 32);
    mbar_init(&zt_ready[0], 2 * TC_WORKERS / 32);
    mbar_init(&zt_ready[1], 2 * TC_WORKERS / 32);
    mbar_init(mma_done_b, 1);
    mbar_init(rb_free, 2 * TC_WORKERS / 32);
    mbar_fence_init();
  }
  if (tid < PINN_NSUMS) red[tid] = 0.0;
  for (int i = tid; i < TC_MAX_HH * TC_H; i += TC_THREADS) db_s[i] = 0.f;
  for (int i = tid; i < TC_H * 4; i += TC_THREADS) w0s[i] = __ldg(w0p + (i >> 2) * 8 + (i & 3));
  if (warp == TC_WORKERS / 32 + 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= TC_WORKERS / 32) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_CONTROL_REGS));
  // (warps 18, 19: spare warps of the control warpgroup -- setmaxnreg works on whole warpgroups)
  if (warp == TC_WORKERS / 32) {
    // =========================================== producer ===========================================
    // Both CTAs run the same load sequence; each streams ITS half of every operand:
    //   forward / adjoint job: 8 stages = the 128-column half `rank` of the weight image (32 K-features per stage)
    //   weight-gradient job:   16 stages = for tile 0 then tile 1 of the pair, per 16 rows the piece `rank` of Zbar_l (A
    //                          operand, its M half) and of a_{l-1} (B operand, its N half): two 8 KB copies on one barrier
    // ONE LANE PER RING STAGE: a thread needs ~400 cycles (1,400 while the workers' spill stores and REDs fill the LSU
    // queue) to get through wait -> expect_tx -> cp.async.bulk, whatever the copy's size (tools/mma_interf_probe.cu), so a
    // single producer thread cannot refill faster than one stage per 400 .. 1,400 cycles; lane s owns stage s and the items
    // s, s + S, s + 2S, ... of the load sequence, so up to S copies are being issued at once.
    if (lane < TC_STAGES) {
#ifndef KO_RING
      const size_t half_off = (size_t)rank * (TC_H * TC_H / 2);
      const int n_fwd = NHH * TC_WCHUNKS * XP;                      // items of the forward jobs
      const int n_rev = TC_WCHUNKS * XP + 16;                       // items per reverse layer: adjoint job, weight-gradient job
      const int per_tile = n_fwd + XP + (BWD ? 1 + NHH * n_rev : 0);
      const long long total = (long long)my_tiles * per_tile;
      unsigned char* stage = ring + lane * TC_STAGE_BYTES;
      uint32_t pph = 1;     // parity to wait for on empty[lane] (first pass: fresh barriers count as released)
      long long it = (long long)lane / per_tile;
      int r = lane - (int)it * per_tile;                            // item r of tile pair it
      for (long long k = lane; k < total; k += TC_STAGES, pph ^= 1u) {
        const float* src = nullptr;
        const float* src2 = nullptr;   // weight-gradient items: second copy (the a piece)
        uint32_t bytes = TC_STAGE_BYTES;
        if (r < n_fwd) {                                            // forward job of layer hl + 1, chunk c, part x (hi / lo)
          const int hl = r / (TC_WCHUNKS * XP), q = r - hl * (TC_WCHUNKS * XP);
          const int c = q / XP, x = q - c * XP;
          src = A.packed + (size_t)hl * LSTRIDE + (size_t)x * LO_OFF + half_off + (size_t)c * TC_STAGE_FLOATS;
        } else if (r < n_fwd + XP) {                                // last layer, forward: 256 x 16 B image of this CTA
          src = edge + TC_EDGE_E1 + (size_t)(2 * rank + (r - n_fwd)) * TC_CHUNK16;
        } else if (r == n_fwd + XP) {                               // last layer, reverse: 8 x 128 B image(s) of this CTA
          src = edge + TC_EDGE_E2 + (size_t)rank * 2048;
          bytes = 4096 * XP;
        } else {
          const int q = r - (n_fwd + XP + 1);
          const int li = q / n_rev, w = q - li * n_rev;
          const int l = L - 2 - li;
          if (w < TC_WCHUNKS * XP) {                                // adjoint job of layer l
            const int c = w / XP, x = w - c * XP;
            src = A.packed + (size_t)(l - 1) * LSTRIDE + (size_t)x * LO_OFF + (size_t)TC_H * TC_H + half_off +
                  (size_t)c * TC_STAGE_FLOATS;
          } else {                                                  // weight-gradient job of layer l: tile t, 16-row chunk rr
            const int wq = w - TC_WCHUNKS * XP, t = wq >> 3, rr = wq & 7;
            const long long nzt = it * NHH + li;
            mbar_wait(slab_ready, (uint32_t)(it & 1));              // both tiles' activation spills are written and fenced
            mbar_wait(&zt_ready[nzt & 1], (uint32_t)((nzt >> 1) & 1));  // Zbar_l of both tiles has been spilled
            src = slab_pair[t] + (size_t)(NHH + (l & 1)) * TC_IMG + (size_t)(2 * rr + rank) * TC_PIECE;
            src2 = slab_pair[t] + (size_t)(l - 1) * TC_IMG + (size_t)(2 * rr + rank) * TC_PIECE;
          }
        }
#ifdef KO_HALFW
        if (!src2) bytes >>= 1;   // timing experiment: half of every weight copy (what a 4-CTA multicast would load per CTA)
#endif
#ifdef TC_TMAP_RING
        mbar_wait(&empty[lane], pph);
        if (rank == 0) mbar_expect_tx(&full[lane], 2 * bytes);   // both CTAs' copies complete on the leader's barrier
        {
          const uint32_t lf = mapa_u32(&full[lane], 0);
          if (src2) {
            tma_load_2d_pair(stage, &tm_p, (int)((src - A.slab) >> 8), lf);
            tma_load_2d_pair(stage + TC_STAGE_BYTES / 2, &tm_p, (int)((src2 - A.slab) >> 8), lf);
          } else if (bytes == TC_STAGE_BYTES) {
            tma_load_2d_pair(stage, &tm_w, (int)((src - A.packed) >> 8), lf);
          } else {
            tma_load_2d_pair(stage, &tm_rl, (int)((src - A.packed) >> 8), lf);
          }
        }
        r += TC_STAGES;
        while (r >= per_tile) r -= per_tile, ++it;
        continue;
#endif
        mbar_wait(&empty[lane], pph);
        mbar_expect_tx(&full[lane], bytes);
        if (src2) {
#ifdef KO_DW1COPY
          tma_load_1d(stage, src, TC_STAGE_BYTES, &full[lane]);   // timing experiment: one 16 KB copy per weight-gradient stage
#else
          tma_load_1d(stage, src, TC_STAGE_BYTES / 2, &full[lane]);
          tma_load_1d(stage + TC_STAGE_BYTES / 2, src2, TC_STAGE_BYTES / 2, &full[lane]);
#endif
        } else {
          tma_load_1d(stage, src, bytes, &full[lane]);
        }
        r += TC_STAGES;
        while (r >= per_tile) r -= per_tile, ++it;
      }
#endif
    }
  } else if (warp == TC_WORKERS / 32 + 1) {
    // =========================================== MMA issuer =========================================
#if defined(KO_RING) || defined(TC_TMAP_RING)
    if (false) {
#elif defined(RELAY_ONE)
    if (lane == 0 && rank != 0) {
      // follower: relay "my stage has landed" to the leader's issuer (1-D bulk copies cannot signal a peer barrier).
      // ONE lane, stages in ring order (the order the leader consumes them in): several lanes of one warp parked in
      // mbarrier.try_wait on different stages delay each other.
      const int per_tile = NHH * TC_WCHUNKS * XP + XP + (BWD ? 1 + NHH * (TC_WCHUNKS * XP + 16) : 0);
      const long long total = (long long)my_tiles * per_tile;
#ifdef SPLIT_FULL
      const uint32_t fp0 = mapa_u32(&full_peer[0], 0);
#else
      const uint32_t fp0 = mapa_u32(&full[0], 0);
#endif
      uint32_t par = 0, st = 0;
      for (long long c = 0; c < total; ++c) {
        mbar_wait(&full[st], par);
        mbar_arrive_cluster_relaxed(fp0 + 8u * st);
        if (++st == TC_STAGES) st = 0, par ^= 1u;
      }
    }
    if (false) {
#else
    if (lane < TC_STAGES && rank != 0) {
#endif
      // follower: relay "my stage has landed" to the leader's issuer (1-D bulk copies cannot signal a peer barrier);
      // one lane per ring stage so the hand-offs of different stages overlap
      const int per_tile = NHH * TC_WCHUNKS * XP + XP + (BWD ? 1 + NHH * (TC_WCHUNKS * XP + 16) : 0);
      const long long total = (long long)my_tiles * per_tile;
      const long long passes = (total - lane + TC_STAGES - 1) / TC_STAGES;
#ifdef SPLIT_FULL
      const uint32_t fp = mapa_u32(&full_peer[lane], 0);
#else
      const uint32_t fp = mapa_u32(&full[lane], 0);
#endif
      uint32_t par = 0;
      for (long long c = 0; c < passes; ++c) {
        mbar_wait(&full[lane], par);
        mbar_arrive_cluster_relaxed(fp);
        par ^= 1u;
      }
    }
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc_k = umma_idesc(2 * TC_M, TC_H, 0, 0);
      constexpr uint32_t idesc_mn = umma_idesc(2 * TC_M, TC_H, 1, 1);
      constexpr uint32_t idesc_last = umma_idesc(2 * TC_M, 32, 0, 0);   // 256 -> o forward: N = 2 x 16 (zero-padded)
      // The issue loop runs on ONE thread: every scalar instruction in it is on the tensor pipe's critical path
      // (a 128x256x8 TF32 MMA retires in ~130 cycles, tools/mma_rate_probe.cu).  Descriptors are therefore
      // built once; per MMA only the 14-bit start-address field (16-byte units) advances by a compile-time
      // constant; the ring stage index is a small running counter.
      const uint64_t ad_op = umma_desc(smem_u32(op), OP_LBO, 128);            // + kstep * (2 * OP_LBO / 16)
      const uint64_t bd_k = umma_desc(smem_u32(ring), (TC_H / 2) * 16, 128);  // K-major half-width weight chunk in stage 0; also the
                                                                              // [2][128 rows][4] reverse last-layer image
      const uint64_t d_mn = umma_desc_mn(smem_u32(ring), 2048, 512);          // MN-major spill piece in stage 0
      const uint64_t bd_last = umma_desc(smem_u32(ring), 16 * 16, 128);       // [k/4][16 rows][4] last-layer image in stage 0
      const uint64_t ad_outs = umma_desc(smem_u32(outs), TC_M * 16, 128);     // [2][128 rows][4] adjoint seeds (K = 8 outputs)
      const uint32_t hi_k = (uint32_t)(ad_op >> 32), hi_mn = (uint32_t)(d_mn >> 32);   // (all K-major descriptors share SBO = 128)
      const uint32_t op_lo = (uint32_t)ad_op, ring_lo = (uint32_t)bd_k, last_lo = (uint32_t)bd_last, outs_lo = (uint32_t)ad_outs;
      constexpr uint32_t STG = TC_STAGE_BYTES / 16;
      uint32_t rp = 0;      // parity of the ring pass (flips every TC_STAGES chunks)
      uint32_t rs = 0;      // ring stage of the next chunk
      int jobs = 0, nB = 0;
#ifdef PINN_TC_PHASES
      long long iw_ready = 0, iw_full = 0, iw_peer = 0, iw_rb = 0, i_gemm = 0, i_dw = 0, iw_full_g = 0, iw_peer_g = 0;
#define ITM(acc, stmt) { const long long a_ = clock64(); stmt; acc += clock64() - a_; }
#else
#define ITM(acc, stmt) stmt;
#endif
      // every epilogue hands the operand image over in four 64-feature slices (= two ring stages of K each)
      auto wait_slice = [&]() {
        TRS(2640)
        ITM(iw_ready, mbar_wait(&op_ready[jobs & 3], (uint32_t)((jobs >> 2) & 1)))
        TRS(2650)
        ++jobs;
        tc_fence_after();
      };
      // all four slices: every warp signals its slices in order, so the last slice's barrier completing means the other three
      // have completed too -- one poll instead of four (a poll costs hundreds of cycles while the drain's REDs fill the LSU queue)
      auto wait_ready = [&]() {
#ifdef WAIT_ALL_SLICES
        for (int q = 0; q < 4; ++q) wait_slice();
#else
        jobs += 3;
        wait_slice();
#endif
      };
      int nedge = 0;
      // D[256 x 256] = OP (K-major, 256 features; 128 rows in each CTA) * weight half-images (K-major, 128 columns in each CTA)
      // into TMEM columns dcol..dcol+255.  PIPELINED: consume the operand image slice by slice as the epilogue of the
      // previous layer produces it (the accumulator of that layer sits in the OTHER 256 columns, so the epilogue can
      // still be reading it); otherwise all four slices are awaited first.
      auto gemm_k = [&](uint32_t dcol, bool pipelined) {
        TR(2000)
        if (!pipelined) wait_ready();
        TR(2100)
#pragma unroll 1
        for (int c = 0; c < TC_WCHUNKS; ++c) {
          if (pipelined && (c & 1) == 0) wait_slice();
#pragma unroll
          for (int x = 0; x < XP; ++x) {   // split-operand mode: the same operand rows against the hi, then the lo weight chunk
            const uint32_t s = rs;
            TRS(2610)
            RINGWAIT(ITM(iw_full_g, mbar_wait(&full[s], rp)))
            TRS(2620)
#ifdef SPLIT_FULL
            RINGWAIT(ITM(iw_peer_g, mbar_wait(&full_peer[s], rp)))
#endif
            TRS(2600)
            const uint32_t b_lo = ring_lo + s * STG, a_lo = op_lo + (uint32_t)c * (4 * (2 * OP_LBO / 16));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // 8 contraction features per MMA = two 16-byte K chunks
              umma_tf32_lh(tmem_base + dcol, a_lo + (uint32_t)(kk * (2 * OP_LBO / 16)), hi_k, b_lo + (uint32_t)(kk * (2 * (TC_H / 2) * 16 / 16)), hi_k,
                           idesc_k, (c > 0 || kk > 0 || x > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            TRS(2630)
            if (++rs == TC_STAGES) rs = 0, rp ^= 1u;
          }
        }
        umma_commit(mma_done);
        TR(2200)
      };
      for (int it = 0; it < my_tiles; ++it) {
        TR_SET(true)
        for (int hl = 0; hl < NHH; ++hl) {
          // forward job of layer hl+1 into TMEM half hl & 1: layer 1 into columns 0..255, because columns 256..511 may still
          // hold the previous tile's last weight-gradient accumulator (its drain comes after this tile's layer 0)
          if (BWD && hl == 1 && nB > 0) {
            ITM(iw_rb, mbar_wait(rb_free, (uint32_t)((nB - 1) & 1)))
            tc_fence_after();
          }
          ITM(i_gemm, gemm_k((uint32_t)((hl & 1) * 256), true))
        }
        {
          // last layer forward: D[256 x 32] = OP * Wlast^T (columns >= o are zero), one ring stage per image
          wait_ready();
#pragma unroll
          for (int x = 0; x < XP; ++x) {
            const uint32_t s = rs;
            RINGWAIT(ITM(iw_full_g, mbar_wait(&full[s], rp)))
#ifdef SPLIT_FULL
            RINGWAIT(ITM(iw_peer_g, mbar_wait(&full_peer[s], rp)))
#endif
            const uint32_t b_lo = last_lo + s * STG;
#pragma unroll 4
            for (int kstep = 0; kstep < TC_H / 8; ++kstep)
              umma_tf32_lh(tmem_base, op_lo + (uint32_t)(kstep * (2 * OP_LBO / 16)), hi_k, b_lo + (uint32_t)(kstep * (2 * 16 * 16 / 16)), hi_k, idesc_last,
                           (kstep > 0 || x > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            if (++rs == TC_STAGES) rs = 0, rp ^= 1u;
          }
          umma_commit(mma_done);
        }
        if (BWD) {
          {
            // last layer reverse: Abar_{L-2}[256 x 256] = seeds[256 x 8] * Wlast: ONE MMA (K = 8)
            ITM(iw_ready, mbar_wait(edge_ready, (uint32_t)(nedge & 1)))
            ++nedge;
            tc_fence_after();
            const uint32_t s = rs;
            RINGWAIT(ITM(iw_full_g, mbar_wait(&full[s], rp)))
#ifdef SPLIT_FULL
            RINGWAIT(ITM(iw_peer_g, mbar_wait(&full_peer[s], rp)))
#endif
            umma_tf32_lh(tmem_base, outs_lo, hi_k, ring_lo + s * STG, hi_k, idesc_k, 0u);
            if (X3) umma_tf32_lh(tmem_base, outs_lo, hi_k, ring_lo + s * STG + 4096 / 16, hi_k, idesc_k, 1u);   // lo image of W_last
            umma_commit(&empty[s]);
            if (++rs == TC_STAGES) rs = 0, rp ^= 1u;
            umma_commit(mma_done);
          }
          // weight-gradient job of one layer into TMEM columns 256..511: both operands MN-major (contraction over
          // the 2 x 128 rows of the pair's tiles); per stage = 16 rows: [Zbar half-chunk 8 KB][A_in half-chunk 8 KB]
          auto dw_job = [&]() {
            TR(2300)
            if (nB > 0) {
              ITM(iw_rb, mbar_wait(rb_free, (uint32_t)((nB - 1) & 1)))   // the previous accumulator has been drained in both CTAs
              tc_fence_after();
            }
            ++nB;
            TR(2400)
#pragma unroll 1
            for (int q = 0; q < 16; ++q) {
              const uint32_t s = rs;
              TRS(2710)
              RINGWAIT(ITM(iw_full, mbar_wait(&full[s], rp)))
              TRS(2720)
#ifdef SPLIT_FULL
              RINGWAIT(ITM(iw_peer, mbar_wait(&full_peer[s], rp)))
#endif
              TRS(2700)
              // stage = [Zbar piece 8 KB | a piece 8 KB]; the accumulator is the TRANSPOSED weight gradient a^T Zbar (A operand =
              // the a piece, M = input features; B operand = the Zbar piece, N = output features), so that a TMEM lane is one
              // input feature and the drain's warp-wide REDs hit 128 contiguous bytes of a row of dW
              const uint32_t dz = ring_lo + s * STG, da = dz + 512u;
              if (X3) {
                // the two K atoms of a piece are the hi rows and the lo rows of the same jets: hi.hi + lo.hi + hi.lo
                umma_tf32_lh(tmem_base + 256u, da, hi_mn, dz, hi_mn, idesc_mn, q > 0 ? 1u : 0u);
                umma_tf32_lh(tmem_base + 256u, da + 64u, hi_mn, dz, hi_mn, idesc_mn, 1u);
                umma_tf32_lh(tmem_base + 256u, da, hi_mn, dz + 64u, hi_mn, idesc_mn, 1u);
              } else {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                  umma_tf32_lh(tmem_base + 256u, da + (uint32_t)(kk * 64), hi_mn, dz + (uint32_t)(kk * 64), hi_mn, idesc_mn,
                               (q > 0 || kk > 0) ? 1u : 0u);
              }
              umma_commit(&empty[s]);
              TRS(2730)
              if (++rs == TC_STAGES) rs = 0, rp ^= 1u;
            }
            umma_commit(mma_done_b);
            TR(2500)
          };
          for (int l = L - 2; l >= 1; --l) {
            ITM(i_gemm, gemm_k(0u, false))     // adjoint of the layer input, once Zbar_l is complete in both operand images
            ITM(i_dw, dw_job())                // weight gradient of layer l
          }
        }
      }
#ifdef PINN_TC_PHASES
      if (blockIdx.x == 0)
        printf("TC issuer (cycles per tile pair): wait op_ready %lld | fwd+adj jobs %lld (14) | dW jobs %lld (7) | gemm waits: full %lld, peer %lld | dW waits: full %lld, peer %lld, rb_free %lld\n",
               iw_ready / my_tiles, i_gemm / my_tiles, i_dw / my_tiles, iw_full_g / my_tiles, iw_peer_g / my_tiles, iw_full / my_tiles, iw_peer / my_tiles, iw_rb / my_tiles);
#endif
    }
  }
  } else {
    // =========================================== workers ============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_WORKER_REGS));
    const int sp = warp & 3, half = warp >> 2;   // the four warps of a subpartition split every 64-feature slice
    const int cbase = half * TC_WCOLS;           // (drain only: contiguous column range of this warp)
    const int pp = lane >> 2, cq = lane & 3;     // point within the subpartition, column pair within 8 columns
    const int pt = X3 ? sp * 4 + (pp & 3) : sp * 8 + pp;   // this thread's point of the tile
    const int mrow0 = sp * 32 + pp;              // row of its slot 0; slot a sits at row mrow0 + 8 a (TF32 mode: slot = jet;
                                                 // split-operand mode: slot 2 jj + h = part h of jet 2 jh + jj, jh = pp / 4)
    const bool lead = lane < 16;                 // split-operand mode: this lane holds jets 0, 1 (the partner lane ^ 16: jets 2, 3)
    const uint32_t tmem_sp = tmem_base + ((uint32_t)(sp * 32) << 16);
    const float inv_cnt = (kind == PINN_RES_CONT_ONLY && A.mask_count) ? 1.0f / *A.mask_count : 0.f;
    // This thread's features in block b = slice b of the layer (64 features), 16 of which belong to this warp.  The rows
    // of the weight images are permuted inside every group of 16 (pack_tc_kernel) so that MMA column 16 g + 8 u + 2 cq + e
    // holds feature 16 g + 4 cq + 2 u + e: the four values a thread gets from a 16x256b.x2 load, v[j][i = 2u+e], are the
    // FOUR CONSECUTIVE features
    //   f = 64 b + 16 half + 4 cq + i
    // i.e. one 16-byte unit of the operand image and of the spill image (128-bit shared / global accesses).  Slices are
    // finished by all warps at the same time, so the next layer's MMA can start on slice 0 while slice 1 is computed.
    // operand image: (jet j, block b) at op_thr + 16 b * OP_LBO + j * 128
    unsigned char* op_thr = op + (4 * half + cq) * OP_LBO + mrow0 * 16;
    // spill image [16-row chunk (8)][feature half (2)][piece]: (j, b) at float
    //   img_thr + (j>>1)*4096 + (j&1)*256 + (b>>1)*2048 + (2(b&1) + half/2)*512 + ((2(half&1) + cq/2) ^ (pp&3))*8
    const int img_thr = (2 * sp) * (2 * TC_PIECE) + (half >> 1) * 512 + pp * 32 + (cq & 1) * 4;
    const int img_swz = 2 * (half & 1) + (cq >> 1);
    auto img_off = [&](int j, int b) {
      return img_thr + (j >> 1) * (2 * TC_PIECE) + (j & 1) * 256 + (b >> 1) * TC_PIECE + (2 * (b & 1)) * 512 +
             ((img_swz ^ (pp & 3)) << 3);
    };
    // slab of this CTA: a_0 .. a_{L-3} (image l = a_l, the B operand of the weight-gradient job of layer l+1), then two
    // Zbar buffers used alternately (Zbar_l in buffer l & 1): a Zbar image is dead as soon as its weight-gradient job has
    // read it, so re-using two buffers keeps it in L2 instead of writing every layer's copy back to HBM
    auto zbuf = [&](int l) { return slab + (size_t)(NHH + (l & 1)) * TC_IMG; };
    int nzs = 0;  // Zbar spills published
    int mj = 0;   // adjoint / forward MMA jobs waited for
    int nbw = 0;  // weight-gradient half jobs drained
    TCT_DECL
    auto wait_mma = [&]() {
      mbar_wait(mma_done, (uint32_t)(mj & 1));
      ++mj;
      tc_fence_after();
    };
    // cross-CTA signalling: every thread fences, one elected lane per warp arrives
    const uint32_t op_ready_leader = mapa_u32(op_ready, 0), rb_free_leader = mapa_u32(rb_free, 0);
    int nsl = 0;  // slices handed over (selects the barrier)
    // The operand image is read by this SM's own tensor core only, so a CTA-scope proxy fence per thread is enough;
    // the arrive itself is relaxed: a releasing arrive would first wait for this thread's spill stores to reach L2.
    auto signal_slice = [&]() {   // this warp's part of one 64-feature slice is written (and its TMEM reads are done)
      tc_fence_before();
#ifndef KO_FENCE
      fence_async_proxy_smem();
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(op_ready_leader + 8u * (uint32_t)(nsl & 3));
      ++nsl;
    };
    const uint32_t edge_ready_leader = mapa_u32(edge_ready, 0);
    auto signal_edge = [&]() {
      tc_fence_before();
      fence_async_proxy_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(edge_ready_leader);
    };
    auto publish_spill = [&](uint64_t* bar) {   // this thread's spill stores -> visible to both CTAs' TMA engines at L2
      __threadfence();          // every thread: its own stores are performed at GPU scope ...
      fence_async_proxy();      // ... and ordered before async-proxy (TMA) reads
      __syncwarp();
      if (lane == 0) {          // the fences above did the ordering; the arrives are plain hand-offs
        mbar_arrive_cluster_relaxed(mapa_u32(bar, 0));
        mbar_arrive_cluster_relaxed(mapa_u32(bar, 1));
      }
    };
    // all four jets of this thread's point for the 4 features of block b: v[j][2u+e]
    auto ld_block = [&](int b, float (&v)[4][4], uint32_t buf = 0u) {
      uint32_t ra[8], rb[8];
      const uint32_t col = buf * 256u + (uint32_t)(64 * b + 16 * half);
      tmem_ld_16x256b_x2_nowait(tmem_sp + col, ra);
      tmem_ld_16x256b_x2_nowait(tmem_sp + (16u << 16) + col, rb);
      tmem_wait_ld();
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          v[0][2 * u + e] = __uint_as_float(ra[4 * u + e]);
          v[1][2 * u + e] = __uint_as_float(ra[4 * u + 2 + e]);
          v[2][2 * u + e] = __uint_as_float(rb[4 * u + e]);
          v[3][2 * u + e] = __uint_as_float(rb[4 * u + 2 + e]);
        }
    };
    auto st_op_block = [&](int b, const float (&v)[4][4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(op_thr + 16 * b * OP_LBO + j * 128) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
    };
    auto ld_op_block = [&](int b, float (&v)[4][4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(op_thr + 16 * b * OP_LBO + j * 128);
        v[j][0] = t.x, v[j][1] = t.y, v[j][2] = t.z, v[j][3] = t.w;
      }
    };
    auto st_img_block = [&](float* img, int b, const float (&v)[4][4]) {
#ifndef KO_SPILL
#pragma unroll
      for (int j = 0; j < 4; ++j) st_global_v4(img + img_off(j, b), v[j][0], v[j][1], v[j][2], v[j][3]);
#endif
    };
    // Zbar spill: the two buffers are rewritten every other layer and read once in between -- keep them in L2 (evict_last)
    // so that the activation images streaming through the cache do not push them out to HBM
#ifndef NO_ZBAR_HINT
    const uint64_t pol_z = l2_policy_evict_last();
#endif
    auto st_zimg_block = [&](float* img, int b, const float (&v)[4][4]) {
#ifndef KO_SPILL
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#ifndef NO_ZBAR_HINT
        st_global_v4_hint(img + img_off(j, b), v[j][0], v[j][1], v[j][2], v[j][3], pol_z);
#else
        st_global_v4(img + img_off(j, b), v[j][0], v[j][1], v[j][2], v[j][3]);
#endif
      }
#endif
    };
    auto ld_img_block = [&](const float* img, int b, float (&v)[4][4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#ifdef KO_RELOAD
        const float4 t = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
#else
        const float4 t = ld_global_v4(img + img_off(j, b));
#endif
        v[j][0] = t.x, v[j][1] = t.y, v[j][2] = t.z, v[j][3] = t.w;
      }
    };
    // forward activation: pre-activation jets z (value row already carries the bias) -> post-activation jets
    auto activate = [&](float (&z)[4][4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = tanh_approx(z[0][i]);
        const float s = fmaf(-a, a, 1.f);
        z[0][i] = round_tf32(a);
        z[1][i] = round_tf32(s * z[1][i]);
        z[2][i] = round_tf32(s * z[2][i]);
        z[3][i] = round_tf32(s * z[3][i]);
      }
    };
    // adjoint through the activation: ab = adjoint of the post-activation jets, act = stored post-activation
    // jets -> ab = adjoint of the pre-activation jets (Zbar)
    auto adjoint = [&](float (&ab)[4][4], const float (&act)[4][4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = act[0][i];
        const float s = fmaf(-a, a, 1.f);
        const float pr = fmaf(ab[1][i], act[1][i], fmaf(ab[2][i], act[2][i], ab[3][i] * act[3][i]));
        ab[0][i] = round_tf32(fmaf(-2.f * a, pr, ab[0][i] * s));
        ab[1][i] = round_tf32(ab[1][i] * s);
        ab[2][i] = round_tf32(ab[2][i] * s);
        ab[3][i] = round_tf32(ab[3][i] * s);
      }
    };
    // ---- split-operand mode ----
    // x -> (hi, lo) with hi = tf32(x), lo = tf32(x - hi): written to the thread's slots 2 jj (hi) and 2 jj + 1 (lo)
    auto split_store = [&](float (&v)[4][4], int i, float x0, float x1) {
      const float h0 = round_tf32(x0), h1 = round_tf32(x1);
      v[0][i] = h0, v[1][i] = round_tf32(x0 - h0);
      v[2][i] = h1, v[3][i] = round_tf32(x1 - h1);
    };
    // v: the thread's four accumulator slots of block b (hi and lo rows of its two jets) -> post-activation jets, split.
    // bias: the layer's bias for the thread's features (added to the value jet, which the lead lane holds).
    // Each lane evaluates tanh for two of the four features; the pair exchanges pre-activations, s = 1 - a^2 and a.
    auto activate_x3 = [&](float (&v)[4][4], const float (&bias)[4]) {
      float z0[4], z1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z0[i] = v[0][i] + v[1][i] + (lead ? bias[i] : 0.f);   // lead: value pre-activation; partner: tangent jet 2
        z1[i] = v[2][i] + v[3][i];                            // lead: tangent jet 1;         partner: tangent jet 3
      }
      const float r2 = __shfl_xor_sync(0xffffffffu, z0[2], 16), r3 = __shfl_xor_sync(0xffffffffu, z0[3], 16);
      const float ta = tanhf(lead ? z0[0] : r2), tb = tanhf(lead ? z0[1] : r3);   // lead: features 0, 1; partner: 2, 3
      const float sa = fmaf(-ta, ta, 1.f), sb = fmaf(-tb, tb, 1.f);
      const float osa = __shfl_xor_sync(0xffffffffu, sa, 16), osb = __shfl_xor_sync(0xffffffffu, sb, 16);
      const float ota = __shfl_xor_sync(0xffffffffu, ta, 16), otb = __shfl_xor_sync(0xffffffffu, tb, 16);
      const float s[4] = {lead ? sa : osa, lead ? sb : osb, lead ? osa : sa, lead ? osb : sb};
      const float a[4] = {ta, tb, ota, otb};   // (meaningful in the lead lane only)
#pragma unroll
      for (int i = 0; i < 4; ++i) split_store(v, i, lead ? a[i] : s[i] * z0[i], s[i] * z1[i]);
    };
    // ab: accumulator slots of the adjoint job (adjoint of the post-activation jets), act: the stored post-activation
    // jets (hi / lo slots) -> ab = Zbar, split; zb0 = the un-split value-row Zbar (lead lanes; for the bias gradient)
    auto adjoint_x3 = [&](float (&ab)[4][4], const float (&act)[4][4], float (&zb0)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g0 = ab[0][i] + ab[1][i], g1 = ab[2][i] + ab[3][i];
        const float c0 = act[0][i] + act[1][i], c1 = act[2][i] + act[3][i];
        const float prp = lead ? g1 * c1 : fmaf(g0, c0, g1 * c1);   // this lane's part of sum_{j>=1} abar_j * adot'_j
        const float ex = __shfl_xor_sync(0xffffffffu, lead ? c0 : prp, 16);   // lead receives the partner's part, the partner a'
        const float a = lead ? c0 : ex;
        const float sgm = fmaf(-a, a, 1.f);
        const float x0 = lead ? fmaf(-2.f * a, prp + ex, g0 * sgm) : g0 * sgm;
        zb0[i] = x0;
        split_store(ab, i, x0, g1 * sgm);
      }
    };
    // bias gradient, split-operand mode: the value-row Zbar of the warp's 4 points sits in lanes 4 q + cq (< 16):
    // transposing butterfly over lane bits 3, 2, then one shared-memory atomic per feature (16 lanes)
    auto db_block_x3 = [&](float* dbl, int b, const float (&zb)[4]) {
      const bool b3 = lane & 8, b2 = lane & 4;
      float k0 = b3 ? zb[2] : zb[0], k1 = b3 ? zb[3] : zb[1];
      const float s0 = b3 ? zb[0] : zb[2], s1 = b3 ? zb[1] : zb[3];
      k0 += __shfl_xor_sync(0xffffffffu, s0, 8);
      k1 += __shfl_xor_sync(0xffffffffu, s1, 8);    // holds features 2 b3 + {0, 1}
      float k = b2 ? k1 : k0;
      const float sx = b2 ? k0 : k1;
      k += __shfl_xor_sync(0xffffffffu, sx, 4);     // holds feature 2 b3 + b2, summed over the 4 points
      if (lead) atomicAdd(dbl + 64 * b + 16 * half + 4 * cq + 2 * (b3 ? 1 : 0) + (b2 ? 1 : 0), k);
    };
    // bias gradient of a hidden layer: sum over the tile's points of the value-row Zbar.  The 8 points of a
    // warp sit in lanes 4 pp + cq: halving butterfly over lane bits 4, 3, 2, then one shared-memory atomic
    // per feature (16 lanes), accumulated over all tiles and flushed once at kernel exit.
    auto db_block = [&](float* dbl, int b, const float (&zb)[4]) {
      const bool b4 = lane & 16, b3 = lane & 8;
      float k0 = b4 ? zb[2] : zb[0], k1 = b4 ? zb[3] : zb[1];
      const float s0 = b4 ? zb[0] : zb[2], s1 = b4 ? zb[1] : zb[3];
      k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
      k1 += __shfl_xor_sync(0xffffffffu, s1, 16);   // holds u = b4, e = 0 / 1
      float k = b3 ? k1 : k0;
      const float s = b3 ? k0 : k1;
      k += __shfl_xor_sync(0xffffffffu, s, 8);      // holds u = b4, e = b3
      k += __shfl_xor_sync(0xffffffffu, k, 4);
      if (!(lane & 4)) atomicAdd(dbl + 64 * b + 16 * half + 4 * cq + 2 * (b4 ? 1 : 0) + (b3 ? 1 : 0), k);
    };

    // Edge-layer gradients (dW0 columns 0..3, db0, dW_last rows 0..3) of this thread's feature, accumulated over ALL tiles of
    // the CTA in registers and flushed once at kernel exit: per tile they were ~9 scalar atomics per thread onto 2,300
    // addresses shared by all 148 CTAs -- same-address REDs that kept the LSU queue full into the next tile's layer 0
    float e0[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, eL[4] = {0.f, 0.f, 0.f, 0.f};
    // input tile + layer 0 (d -> 256) on the FP32 pipes -> operand image, slice by slice
    auto layer0 = [&](int itn) {
      const long long tile_n = 2ll * (pair + (long long)itn * n_pairs) + rank;   // may lie past the end: an all-padding tile
      const long long p0n = tile_n * TP;
      for (int i = tid; i < TP * 8; i += TC_WORKERS) {
        const int pq = i >> 3, c = i & 7;
        const long long gp = p0n + pq;
        xin[i] = (c < d && gp < A.n_points) ? A.inputs[gp * d + c] : 0.f;
      }
      worker_bar();
      {
        float x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = xin[pt * 8 + c];
        const float* b0 = A.params + (long long)d * TC_H;
        for (int b = 0; b < TC_NBLK; ++b) {
          float z[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = 64 * b + 16 * half + 4 * cq + i;
            // columns 0..3 of W0 from shared memory (the 28 KB of L1 beside it do not keep the packed copy across a tile);
            // wider inputs fetch the rest from the packed copy
            const float4 wa = *reinterpret_cast<const float4*>(w0s + f * 4);
            const float4 wb = d > 4 ? __ldg(reinterpret_cast<const float4*>(w0p + f * 8 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            float acc = __ldg(b0 + f);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc = fmaf(x[c], w[c], acc);
            z[0][i] = acc;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
              float t = 0.f;
              if (jj < D.n_dirs) {
                const int dc = D.dir_cols[jj];
#pragma unroll
                for (int c = 0; c < 8; ++c) t = (c == dc) ? w[c] : t;
              }
              z[1 + jj][i] = t;
            }
          }
          if (X3) {
            // both lanes of a pair evaluate the (cheap) layer-0 activation; each keeps its two jets, split into hi / lo
            float v[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = tanhf(z[0][i]);
              const float sg = fmaf(-a, a, 1.f);
              split_store(v, i, lead ? a : sg * z[2][i], lead ? sg * z[1][i] : sg * z[3][i]);
            }
            st_op_block(b, v);
          } else {
            activate(z);
            st_op_block(b, z);
          }
          signal_slice();
        }
      }
    };
    // Tile boundary: the NEXT tile's layer 0 is computed at the end of a tile -- after the layer-0 gradients, before the
    // drain of layer 1 -- so the first forward job of the next tile is ready when the last weight-gradient job leaves the
    // tensor pipe; otherwise the pipe idles through drain + layer-0 gradients + input fetch + layer 0 (~16k cycles per pair).
    bool l0_ready = false;
    for (int it = 0; it < my_tiles; ++it) {
      TR_SET(tid == 0)
      TR(1900)
      const long long tile = 2ll * (pair + (long long)it * n_pairs) + rank;   // may lie past the end: an all-padding tile
      const long long p0 = tile * TP;
      if (!l0_ready) layer0(it);
      l0_ready = false;
      TCT(0)
      {
        // the spill of a_0 is re-read from the operand image AFTER the last slice has been handed over: its stores stall
        // on the LSU queue, and the tensor core should already be running layer 1 while they drain
        if (BWD && NHH >= 1) {
#pragma unroll
          for (int b = 0; b < TC_NBLK; ++b) {
            float z[4][4];
            ld_op_block(b, z);
            st_img_block(slab, b, z);
          }
        }
      }
      TCT(1)
      // ---------------- hidden layers 1..L-2 on the tensor cores ----------------
      for (int l = 1; l <= L - 2; ++l) {
        const float* bias_l = A.params + P0 + (long long)(l - 1) * PH + (long long)TC_H * TC_H + 16 * half + 4 * cq;
        float bl[TC_NBLK][4];
#pragma unroll
        for (int b = 0; b < TC_NBLK; ++b)
#pragma unroll
          for (int i = 0; i < 4; ++i) bl[b][i] = __ldg(bias_l + 64 * b + i);
        TR(1950 + l)
        wait_mma();
        TR(1000 + l)
        TCT(2)
        float* img = slab + (size_t)l * TC_IMG;
        const bool spill = BWD && l <= L - 3;
#pragma unroll
        for (int b = 0; b < TC_NBLK; ++b) {
          float z[4][4];
          ld_block(b, z, (uint32_t)((l & 1) ^ 1));   // forward accumulators alternate between the two TMEM halves (layer 1: columns 0..255)
          if (X3) {
            activate_x3(z, bl[b]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) z[0][i] += bl[b][i];
            activate(z);
          }
          st_op_block(b, z);
          signal_slice();   // slice b of a_l is in the operand image: the next job (layer l+1, or 256 -> o) may consume it
        }
        TR(1100 + l)
        if (spill) {        // spill from the operand image once all slices are handed over (see layer 0)
#pragma unroll
          for (int b = 0; b < TC_NBLK; ++b) {
            float z[4][4];
            ld_op_block(b, z);
            st_img_block(img, b, z);
          }
        }
        TCT(1)
      }
      TR(1190)
      if (BWD) publish_spill(slab_ready);
      TR(1200)
      TCT(3)
      // ---------------- last layer (256 -> o): tensor-core job with N = 32, read back thread-per-row ----------------
      wait_mma();
      TR(1210)
      if (half == 0) {
        const int mm = sp * 32 + lane;   // this thread's TMEM lane = tile row
        float v[8];
        tmem_ld_32x32b_x8(tmem_sp, v);
        if (X3) {                        // hi row + lo row of a jet (8 rows apart); both rows keep the sum
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], 8);
        }
        if (X3 ? ((lane >> 3) == 0 && (lane & 4) == 0) : (((mm >> 3) & 3) == 0)) {      // value rows carry the bias
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < o) v[c] += __ldg(A.params + poffL + (long long)TC_H * o + c);
        }
        *reinterpret_cast<float4*>(outs + mm * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(outs + 512 + mm * 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      tc_fence_before();
      worker_bar();
      TR(1211)
      TCT(13)
      // ---------------- residual / misfit epilogue: warp 0, one lane per point ----------------
      if (warp == 0) {
        const bool isp = lane < TP;
        const int pl = isp ? lane : 0;
        const long long gp = p0 + pl;
        TileJets<X3> acc{outs, pl};
        float ls[PINN_NSUMS];
        residual_epilogue<4>(D, acc, isp, isp && gp < A.n_points, gp, xin + pl * 8,
                             EpiArgs{A.targets, nullptr, {nullptr, nullptr, nullptr}, A.out,
                                     {A.dout[0], A.dout[1], A.dout[2]}, A.inv_n_res, A.inv_n_fid, inv_cnt},
                             ls);
        TR(1212)
#pragma unroll
        for (int i = 0; i < PINN_NSUMS; ++i) {
          const float v = warp_sum_tc(ls[i]);
          if (lane == 0 && v != 0.f) red[i] += (double)v;
        }
        TR(1213)
        if (X3 && BWD && isp) {   // adjoint seeds -> hi in the jet's hi row, lo in its lo row (A operand of the reverse last-layer MMA)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = acc.row(j);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float x = outs[outs_idx(r, c)];
              const float hi = round_tf32(x);
              outs[outs_idx(r, c)] = hi;
              outs[outs_idx(r + 8, c)] = round_tf32(x - hi);
            }
          }
        }
      }
      worker_bar();
      TR(1220)
      TCT(4)
      if (!BWD) continue;
      signal_edge();    // the adjoint seeds are in the output image: Abar_{L-2} = seeds * Wlast may start

      // =============================== reverse ===============================
      // ---- last layer: dW_last[c][f] = sum_m zbar[m][c] * A[m][f] (thread per feature), db_last ----
      {
        const int f = tid & (TC_H - 1), part = tid / TC_H;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        const unsigned char* ap = op + (f >> 2) * OP_LBO + (f & 3) * 4;
        for (int mm = part * (TC_M / TC_PARTS); mm < (part + 1) * (TC_M / TC_PARTS); ++mm) {
          if (X3 && (mm & 8)) continue;   // split-operand mode: visit the hi rows, add the lo row 8 further
          float a = *reinterpret_cast<const float*>(ap + mm * 16);
          float4 z0 = *reinterpret_cast<const float4*>(outs + mm * 4);
          float4 z1 = *reinterpret_cast<const float4*>(outs + 512 + mm * 4);
          if (X3) {
            a += *reinterpret_cast<const float*>(ap + (mm + 8) * 16);
            const float4 y0 = *reinterpret_cast<const float4*>(outs + (mm + 8) * 4);
            const float4 y1 = *reinterpret_cast<const float4*>(outs + 512 + (mm + 8) * 4);
            z0.x += y0.x, z0.y += y0.y, z0.z += y0.z, z0.w += y0.w;
            z1.x += y1.x, z1.y += y1.y, z1.z += y1.z, z1.w += y1.w;
          }
          acc[0] = fmaf(z0.x, a, acc[0]), acc[1] = fmaf(z0.y, a, acc[1]);
          acc[2] = fmaf(z0.z, a, acc[2]), acc[3] = fmaf(z0.w, a, acc[3]);
          acc[4] = fmaf(z1.x, a, acc[4]), acc[5] = fmaf(z1.y, a, acc[5]);
          acc[6] = fmaf(z1.z, a, acc[6]), acc[7] = fmaf(z1.w, a, acc[7]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) eL[c] += acc[c];
#pragma unroll
        for (int c = 4; c < 8; ++c)
          if (c < o) atomicAdd(A.grad + poffL + (long long)c * TC_H + f, acc[c]);
        if (tid < o) {
          float s = 0.f;
          for (int pq = 0; pq < TP; ++pq) {
            const int r = tile_row<X3>(pq, 0);
            s += outs[outs_idx(r, tid)];
            if (X3) s += outs[outs_idx(r + 8, tid)];
          }
          atomicAdd(A.grad + poffL + (long long)TC_H * o + tid, s);
        }
      }
      worker_bar();
      TR(1230)
      TCT(14)
      // ---- Abar_{L-2} = seeds * W_last (tensor core), through the activation of layer L-2 -> Zbar_{L-2} in place ----
      {
        float* zdst = zbuf(L - 2);
        // bias gradients of the first TC_MAX_HH hidden->hidden layers are staged in shared memory and flushed once at kernel
        // exit; deeper nets add the rest straight into the flat gradient (one atomic per feature per tile)
        float* dbl = (L - 3 < TC_MAX_HH) ? db_s + (size_t)(L - 3) * TC_H
                                         : A.grad + P0 + (long long)(L - 3) * PH + (long long)TC_H * TC_H;
        wait_mma();
        TR(1300)
#pragma unroll
        for (int b = 0; b < TC_NBLK; ++b) {
          float ab[4][4], act[4][4];
          ld_op_block(b, act);
          ld_block(b, ab);
          float zb0[4];
          if (X3) adjoint_x3(ab, act, zb0);
          else adjoint(ab, act);
          st_op_block(b, ab);
          signal_slice();                   // (the adjoint job of layer L-2 starts once all four slices are in)
          st_zimg_block(zdst, b, ab);
          if (X3) db_block_x3(dbl, b, zb0);
          else db_block(dbl, b, ab[0]);
        }
      }
      TR(1310)
      publish_spill(&zt_ready[nzs++ & 1]);
      TR(1320)
      TCT(5)
      // ---- hidden layers L-2 .. 1 ----
      //   tensor core: adjoint job of layer l (TMEM columns 0..255), then the weight-gradient job of layer l over the
      //                rows of both tiles of the pair (columns 256..511)
      //   workers:     adjoint epilogue of layer l (Zbar_{l-1} -> operand image + spill) | drain of layer l
      // drain TMEM columns 256..511 = the transposed weight gradient of layer l: lane = input feature 128 rank + 32 sp + lane,
      // column = output feature.  32x32b loads (thread = lane): for every output feature the warp adds 32 consecutive floats
      // of one row of dW -- 128 contiguous bytes per RED instruction (tools/red_probe.cu: 7.7k cycles per 128 x 256 drain with
      // every CTA draining at once, the L2 atomic rate; 13.2k for sector-sized v2 REDs spread over 8 rows)
      auto drain = [&](int l) {
        mbar_wait(mma_done_b, (uint32_t)(nbw & 1));
        ++nbw;
        tc_fence_after();
        TR(1700 + l)
        const long long poff = P0 + (long long)(l - 1) * PH;
        float* gcol = A.grad + poff + (long long)cbase * TC_H + (int)rank * 128 + sp * 32 + lane;
        const uint32_t ta = tmem_sp + (uint32_t)(256 + cbase);
#pragma unroll
        for (int cb = 0; cb < TC_WCOLS / 16; ++cb) {
          float v[16];
          tmem_ld16(ta + (uint32_t)(cb * 16), v);
#ifdef KO_RED
          if (v[0] == 1234.5f)
#endif
#pragma unroll
          for (int u = 0; u < 16; ++u)
            red_add_f32(gcol + (long long)(cb * 16 + u) * TC_H, X3 ? v[u] * A.comp_dw : v[u]);
        }
#ifdef ZBAR_DISCARD
        // (-DZBAR_DISCARD, off by default: DRAM writes -16 %, 68 -> 58.5 KB per point in the TF32 mode, but the barrier it
        // needs costs 0.3 .. 1.8 % of the evaluation time and HBM is not what bounds the kernel: profiles/r2d_zbar_discard.log)
        // Zbar_l is dead: every MMA of the pair's weight-gradient job has completed (mma_done_b), so both CTAs' copies of
        // this CTA's buffer have landed.  Drop its lines from L2 without a write-back (the buffer is rewritten in full two
        // layers on): 1,024 lines of 128 bytes, two per worker thread
        {
          const float* zb = zbuf(l);
#pragma unroll
          for (int q = 0; q < TC_IMG * 4 / 128 / TC_WORKERS; ++q)
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(zb + (size_t)(q * TC_WORKERS + tid) * 32) : "memory");
        }
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(rb_free_leader);   // (a releasing arrive would wait for the REDs)
#ifdef ZBAR_DISCARD
        worker_bar();   // no warp may start rewriting this buffer (next layer's epilogue) before every discard is out
#endif
      };
      // the stored activations a_{l-1} the adjoint epilogue of layer l needs are fetched one layer ahead, BEFORE the drain of
      // the previous layer: behind the drain's REDs the loads would sit in the LSU queue until the REDs are through
      float act[TC_NBLK][4][4];
#ifndef RELOAD_LATE
#pragma unroll
      for (int b = 0; b < TC_NBLK; ++b) ld_img_block(slab + (size_t)(L - 3) * TC_IMG, b, act[b]);
#endif
      for (int l = L - 2; l >= 1; --l) {
        // adjoint through the activation of layer l-1 -> Zbar_{l-1} in place (+ spill for its weight gradient)
        {
          float* zdst = zbuf(l - 1);
          const int hh = l >= 2 ? l - 2 : 0;
          float* dbl = (hh < TC_MAX_HH) ? db_s + (size_t)hh * TC_H
                                        : A.grad + P0 + (long long)hh * PH + (long long)TC_H * TC_H;
          const bool hidden = l > 1;
#ifdef RELOAD_LATE
#pragma unroll
          for (int b = 0; b < TC_NBLK; ++b) ld_img_block(slab + (size_t)(l - 1) * TC_IMG, b, act[b]);   // in flight while the adjoint MMA runs
#endif
          TR(1350 + l)
          wait_mma();
          TR(1400 + l)
          TCT(9)
#pragma unroll
          for (int b = 0; b < TC_NBLK; ++b) {
            float ab[4][4];
            ld_block(b, ab);
            float zb0[4];
            if (X3) adjoint_x3(ab, act[b], zb0);
            else adjoint(ab, act[b]);
            st_op_block(b, ab);
            if (hidden) {
              signal_slice();               // (the adjoint job of layer l-1 starts once all four slices are in)
              st_zimg_block(zdst, b, ab);
              if (X3) db_block_x3(dbl, b, zb0);
              else db_block(dbl, b, ab[0]);
            }
          }
        }
        TR(1500 + l)
        TCT(10)
        if (l > 1) publish_spill(&zt_ready[nzs++ & 1]);
        TR(1600 + l)
        TCT(6)
#ifndef RELOAD_LATE
        if (l > 1) {
#pragma unroll
          for (int b = 0; b < TC_NBLK; ++b) ld_img_block(slab + (size_t)(l - 2) * TC_IMG, b, act[b]);   // a_{l-2}, for the next layer
        }
#endif
        if (l > 1) drain(l);   // (layer 1: after the next tile's layer 0, see below)
        TR(1800 + l)
        TCT(7)
      }
      tc_fence_before();
      worker_bar();
      TCT(8)
      // ---- layer 0: dW0[f][c], db0[f] from Zbar_0 (thread per feature) ----
      {
        const int f = tid & (TC_H - 1), part = tid / TC_H;
        const unsigned char* zp = op + (f >> 2) * OP_LBO + (f & 3) * 4;
        float acc[8], tj[3] = {0.f, 0.f, 0.f}, sb = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (int pq = part * (TP / TC_PARTS); pq < (part + 1) * (TP / TC_PARTS); ++pq) {
          auto zrow = [&](int j) {   // Zbar_0 of jet j of point pq for feature f (split-operand mode: hi + lo row)
            const int r = tile_row<X3>(pq, j);
            float v = *reinterpret_cast<const float*>(zp + r * 16);
            if (X3) v += *reinterpret_cast<const float*>(zp + (r + 8) * 16);
            return v;
          };
          const float z0 = zrow(0);
          sb += z0;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaf(z0, xin[pq * 8 + c], acc[c]);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) tj[jj] += zrow(1 + jj);
        }
        for (int jj = 0; jj < D.n_dirs; ++jj) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c == D.dir_cols[jj]) acc[c] += tj[jj];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) e0[c] += acc[c];
        e0[4] += sb;
#pragma unroll
        for (int c = 4; c < 8; ++c)
          if (c < d) atomicAdd(A.grad + (long long)f * d + c, acc[c]);
      }
      worker_bar();
      if (it + 1 < my_tiles) {   // the next tile's input and layer 0: its first forward job can follow the last weight-gradient job
        layer0(it + 1);
        l0_ready = true;
      }
      drain(1);
      TR(1990)
      TCT(12)
    }
    if (BWD) {
      {
        const int f = tid & (TC_H - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < d) atomicAdd(A.grad + (long long)f * d + c, e0[c]);
          if (c < o) atomicAdd(A.grad + poffL + (long long)c * TC_H + f, eL[c]);
        }
        atomicAdd(A.grad + (long long)d * TC_H + f, e0[4]);
      }
      // bias gradients of the hidden layers, accumulated over all tiles of this CTA
      worker_bar();
      for (int i = tid; i < TC_MAX_HH * TC_H; i += TC_WORKERS) {
        const int hl = i / TC_H, f = i - hl * TC_H;
        if (hl < NHH) atomicAdd(A.grad + P0 + (long long)hl * PH + (long long)TC_H * TC_H + f, db_s[i]);
      }
    }
#ifdef PINN_TC_PHASES
    if (blockIdx.x == 0 && tid == 0) {
      long long tot = 0;
      for (int i = 0; i < 15; ++i) tot += tct[i];
      printf("TC phases (cycles per tile, %d tiles): in+bar %lld | L0+fwd-epi %lld | fwd-wait-mma %lld | fence+bar %lld | last+residual %lld | rev-last %lld | fence+signal %lld | drain0 %lld | drain1 %lld | wait-adj %lld | adj-epi %lld | L0-rev %lld | last-layer %lld | dW_last %lld | total %lld\n", my_tiles,
             tct[0] / my_tiles, tct[1] / my_tiles, tct[2] / my_tiles, tct[3] / my_tiles, tct[4] / my_tiles, tct[5] / my_tiles, tct[6] / my_tiles,
             tct[7] / my_tiles, tct[8] / my_tiles, tct[9] / my_tiles, tct[10] / my_tiles, tct[12] / my_tiles, tct[13] / my_tiles, tct[14] / my_tiles, tot / my_tiles);
    }
#endif
  }

  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer may still signal this CTA's barriers / read its shared memory until here
  tc_fence_after();
  if (warp == TC_WORKERS / 32 + 1) tmem_dealloc(tmem_base, 512);
  if (tid < PINN_NSUMS && A.sums && red[tid] != 0.0) atomicAdd(A.sums + tid, red[tid]);
}

// Weight images for the tensor-core jobs, TF32-rounded.  Layer hl (= linear layer hl+1) has four half-width images
// of 128 KB, [direction][half]: each is the B operand one CTA of the pair streams (its 128 of the 256 N columns),
// in 16 KB chunks of 32 contraction features laid out [k/4 (8)][n (128)][k%4] (K-major, no swizzle):
//   forward:  B[n][k] = W[n][k]     half = n / 128   (rows n = output feature, contraction k = input feature)
//   adjoint:  B[n][k] = W[k][n]     half = n / 128   (rows n = input feature,  contraction k = output feature)
// Split-operand mode (x3): w = hi + lo, hi = tf32(w), lo = tf32(w - hi); a layer's block is [fwd hi][adj hi][fwd lo][adj lo].
__global__ void pack_tc_kernel(const __grid_constant__ pinn_desc_t D, const float* __restrict__ params,
                               float* __restrict__ packed, int x3, float comp) {
  const int hl = blockIdx.y;
  const int d = D.widths[0];
  const long long poff = (long long)d * TC_H + TC_H + (long long)hl * ((long long)TC_H * TC_H + TC_H);
  const size_t lstride = (size_t)(x3 ? 4 : 2) * TC_H * TC_H;
  float* fwd = packed + (size_t)hl * lstride;
  float* adj = fwd + (size_t)TC_H * TC_H;
  auto at = [](int n, int k) {   // float offset of B[n][k] inside its direction's two half-images
    // feature n = 16 g + 4 cq + 2 u + e sits in MMA row 16 g + 8 u + 2 cq + e (see the worker addressing)
    const int nr = (n & ~15) | (((n >> 1) & 1) << 3) | (((n >> 2) & 3) << 1) | (n & 1);
    return (size_t)(nr >> 7) * (TC_H * TC_H / 2) + (size_t)(k >> 5) * TC_CHUNK16 + (size_t)((k & 31) >> 2) * 512 +
           (size_t)(nr & 127) * 4 + (size_t)(k & 3);
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TC_H * TC_H; i += gridDim.x * blockDim.x) {
    const int n = i / TC_H, k = i - n * TC_H;
    uint32_t r;
    const float wv = x3 ? params[poff + i] * comp : params[poff + i];   // (comp: see x3_comp)
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(wv));
    const float w = __uint_as_float(r);
    fwd[at(n, k)] = w;
    adj[at(k, n)] = w;
    if (x3) {
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(wv - w));
      const float wl = __uint_as_float(r);
      fwd[2 * TC_H * TC_H + at(n, k)] = wl;
      adj[2 * TC_H * TC_H + at(k, n)] = wl;
    }
  }
  if (hl == 0) {   // zero-padded copies of the two edge layers (read with __ldg by every CTA)
    const int L = D.n_linear, o = D.widths[L];
    const long long poffL = (long long)d * TC_H + TC_H + (long long)(L - 2) * ((long long)TC_H * TC_H + TC_H);
    float* edge = packed + (size_t)(L - 2) * lstride;
    float* w0p = edge + TC_EDGE_W0;
    float* wlp = edge + TC_EDGE_WL;
    float* e1 = edge + TC_EDGE_E1;   // forward last layer:  [rank][hi,lo][k/4 (64)][n (16)][4] = Wlast[n][k] for rank 0, zeros for rank 1
    float* e2 = edge + TC_EDGE_E2;   // reverse last layer:  [rank][hi,lo][k/4 (2)][n (128)][4] = Wlast[k][128 r + n], rows permuted
    auto tf32 = [](float x) {
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
      return __uint_as_float(r);
    };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * TC_CHUNK16; i += gridDim.x * blockDim.x) {
      if (i < TC_H * 8) {
        const int f = i >> 3, c = i & 7;
        w0p[i] = c < d ? params[(long long)f * d + c] : 0.f;
        const int c2 = i / TC_H, f2 = i - c2 * TC_H;
        wlp[i] = c2 < o ? params[poffL + (long long)c2 * TC_H + f2] : 0.f;
        // reverse image: i = c2 (k, 0..7) * 256 + f2 (feature)
        const int n = f2 & 127, nr = (n & ~15) | (((n >> 1) & 1) << 3) | (((n >> 2) & 3) << 1) | (n & 1);
        const float wv = c2 < o ? params[poffL + (long long)c2 * TC_H + f2] : 0.f;
        const float wh = tf32(wv);
        const size_t at2 = (size_t)(f2 >> 7) * 2048 + (c2 >> 2) * 512 + nr * 4 + (c2 & 3);
        e2[at2] = wh;
        e2[at2 + 1024] = tf32(wv - wh);
      }
      // forward image: i < 4096: rank 0, i = k * 16 + n; the rest: rank 1 = zeros
      if (i < TC_CHUNK16) {
        const int k = i >> 4, n = i & 15;
        const float wv = n < o ? params[poffL + (long long)n * TC_H + k] * (x3 ? comp : 1.f) : 0.f;
        const float wh = tf32(wv);
        const size_t at1 = (size_t)(k >> 2) * 64 + n * 4 + (k & 3);
        e1[at1] = wh;
        e1[TC_CHUNK16 + at1] = tf32(wv - wh);
      } else {
        e1[TC_CHUNK16 + i] = 0.f;                           // rank 1, hi
        e1[2 * TC_CHUNK16 + i] = 0.f;                       // rank 1, lo
      }
    }
  }
}

// --------------------------------------------------------------------------------------- host side
constexpr size_t tc_smem_bytes() {
  return (size_t)OP_BYTES + (size_t)TC_STAGES * TC_STAGE_BYTES +
         (size_t)TC_M * 8 * 4 + (size_t)TC_TP * 8 * 4 + (size_t)TC_MAX_HH * TC_H * 4 + PINN_NSUMS * 8 + (3 * TC_STAGES + 11) * 8 + 16 +
         (size_t)TC_H * 4 * 4;
}

// Can this description run on the tensor-core kernel?  (otherwise the caller reports UNSUPPORTED)
bool tc_supported(const pinn_desc_t* D, const char** why) {
  const int L = D->n_linear;
  *why = "";
  if (L < 3) return *why = "needs at least two hidden layers", false;
  for (int i = 1; i < L; ++i)
    if (D->widths[i] != TC_H) return *why = "every hidden layer must be 256 wide", false;
  if (D->activation != PINN_ACT_TANH) return *why = "tanh activation only", false;
  const int k = D->residual_kind;
  if ((k < PINN_RES_CONT_ONLY || k > PINN_RES_WAVE_AVG) && k != PINN_RES_BOUSS_SIMPLE)
    return *why = "needs a PDE residual kind (value-only and external-seed passes use the FP32 kernel)", false;
  return true;
}

int tc_workspace(const pinn_desc_t* D, long long n_points, int sms, size_t* packed_bytes, size_t* slab_bytes,
                 long long* slab_stride, int* grid) {
  const int L = D->n_linear;
  const bool x3 = D->precision == PINN_PREC_TF32X3;
  const int tp = x3 ? TC_TP_X3 : TC_TP;
  long long tiles = (n_points + tp - 1) / tp;
  long long pairs = (tiles + 1) / 2;
  if (pairs > sms / 2) pairs = sms / 2;
  if (pairs < 1) pairs = 1;
  long long g = 2 * pairs;   // CTA pairs (clusters of 2)
  *grid = (int)g;
  *packed_bytes = (size_t)(L - 2) * (x3 ? 4 : 2) * TC_H * TC_H * 4 + (size_t)TC_EDGE_FLOATS * 4;   // + edge-layer block
  *slab_stride = (long long)(L - 2 + 2) * TC_IMG;   // one activation image per hidden layer + two alternating Zbar buffers
  *slab_bytes = (size_t)g * (size_t)(*slab_stride) * 4;
  return PINN_OK;
}

int run_tc_pass(const pinn_desc_t* D, const pinn_eval_args_t* a, bool bwd, void* workspace, size_t ws_bytes,
                cudaStream_t st) {
  int dev = 0, sms = 0;
  PINN_CUDA(cudaGetDevice(&dev));
  PINN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t pk = 0, sl = 0;
  long long stride = 0;
  int grid = 1;
  tc_workspace(D, a->n_points, sms, &pk, &sl, &stride, &grid);
  const size_t pk_al = (pk + 255) & ~size_t(255);
  if (ws_bytes < pk_al + sl) return set_error("workspace too small: %zu < %zu bytes", ws_bytes, pk_al + sl), PINN_E_WORKSPACE;
  if (bwd && ((uintptr_t)a->grad & 15) != 0) return set_error("grad must be 16-byte aligned for the tensor-core path"), PINN_E_ARG;
  float* packed = reinterpret_cast<float*>(workspace);
  float* slab = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + pk_al);
  const bool x3 = D->precision == PINN_PREC_TF32X3;
  if (!(a->flags & PINN_FLAG_SKIP_PACK)) {
    dim3 g(64, D->n_linear - 2);
    pack_tc_kernel<<<g, 256, 0, st>>>(*D, a->params, packed, x3 ? 1 : 0, x3 ? x3_comp(64) : 1.f);
    PINN_CUDA(cudaGetLastError());
  }
  const int tp = x3 ? TC_TP_X3 : TC_TP;
  long long tiles = (a->n_points + tp - 1) / tp;
  if (tiles == 0) return PINN_OK;
  if (tiles > 0x7fffffffLL) return set_error("too many tiles"), PINN_E_UNSUPPORTED;
  TcArgs A;
  A.params = a->params;
  A.packed = packed;
  A.inputs = a->inputs;
  A.targets = D->n_targets > 0 ? a->targets : nullptr;
  A.mask_count = a->mask_count;
  A.grad = a->grad;
  A.sums = a->sums;
  A.out = a->out;
  for (int j = 0; j < PINN_MAX_DIRS; ++j) A.dout[j] = a->dout[j];
  A.slab = slab;
  A.slab_stride = stride;
  A.n_points = a->n_points;
  A.n_tiles = (int)tiles;
  A.inv_n_res = a->n_res_global > 0 ? (float)(1.0 / (double)a->n_res_global) : 0.f;
  A.inv_n_fid = a->n_fid_global > 0 ? (float)(1.0 / (double)a->n_fid_global) : 0.f;
  A.comp_dw = x3 ? x3_comp(48) : 1.f;
  const size_t smem = tc_smem_bytes();
#ifdef TC_TMAP_RING
  // tensor maps over the packed weights (rows of 1 KB; boxes of 16 rows = one stage, and of 4 * XP rows = the reverse
  // last-layer image) and over the spill slab (boxes of 8 rows = one piece)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    PINN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return set_error("cuTensorMapEncodeTiled not available"), PINN_E_UNSUPPORTED;
    encode = (EncodeFn)fn;
  }
  auto make_map = [&](CUtensorMap* tm, void* base, size_t bytes, unsigned box_rows) -> int {
    const cuuint64_t dims[2] = {256, (cuuint64_t)((bytes + 1023) / 1024)};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {256, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, TC_TMAP_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", (int)r), PINN_E_CUDA;
    return PINN_OK;
  };
  alignas(64) CUtensorMap tm_w, tm_rl, tm_p;
  if (int rc = make_map(&tm_w, packed, pk_al, 16)) return rc;
  if (int rc = make_map(&tm_rl, packed, pk_al, x3 ? 8 : 4)) return rc;
  if (int rc = make_map(&tm_p, slab, sl, 8)) return rc;
#endif
  auto go = [&](auto kern) -> int {
    PINN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef TC_TMAP_RING
    kern<<<grid, TC_THREADS, smem, st>>>(*D, A, tm_w, tm_rl, tm_p);
#else
    kern<<<grid, TC_THREADS, smem, st>>>(*D, A);
#endif
    PINN_CUDA(cudaGetLastError());
    return PINN_OK;
  };
  if (x3) return bwd ? go(jet_tc_kernel<true, true>) : go(jet_tc_kernel<false, true>);
  return bwd ? go(jet_tc_kernel<true, false>) : go(jet_tc_kernel<false, false>);
}

}  // namespace pinn

#ifdef PINN_TC_TRACE
extern "C" int pinn_debug_trace(long long* out, int max_events) {
  cudaDeviceSynchronize();
  const int n = max_events < 8192 ? max_events : 8192;
  cudaMemcpyFromSymbol(out, pinn::tc_trace_buf, (size_t)n * 16);
  static long long zeros[2 * 8192];
  cudaMemcpyToSymbol(pinn::tc_trace_buf, zeros, sizeof(zeros));
  return n;
}
#endif
