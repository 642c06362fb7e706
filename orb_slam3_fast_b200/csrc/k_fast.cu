// k_fast.cu — the FAST stage of ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:886-960 serial twin,
// :759-846 TBB) including the arithmetic of cv::FAST(TYPE_9_16, nonmaxSuppression = true) that it calls per cell
// (:923-955 / :810-826).
//
// What the reference does per ~35x35 cell: cv::FAST at iniThFAST; if that yields nothing, again at minThFAST; the
// surviving corners (pixel row-major) are appended to the level's candidate list, cells in row-major order.
// OpenCV's cornerScore is the largest t for which the pixel is still a corner, which has the closed form
//     m = max( v - min_arcs max_{k in arc} p_k ,  max_arcs min_{k in arc} p_k - v ),   score = m - 1,
// over the 16 arcs of 9 consecutive ring pixels, and "corner at threshold T" <=> m > T. So one threshold-free pass
// gives everything both thresholds need (verified against cv2).
//
// Mapping: ONE WARP PER CELL. The cell's (wCell+6)x(hCell+6) raw window is brought into shared memory by ONE TMA
// tile copy per warp (cp.async.bulk.tensor.3d on a per-level {x, y, frame} tensor map, completion on the warp's own
// mbarrier; the box starts at the 16-byte boundary below the window origin, and the level base / pitch / frame stride
// must be 16-byte multiples — otherwise the kernel's LDG staging variant runs) and widened to 16 bit by a
// smem -> smem pass, so that a lane scores 4 horizontally adjacent pixels as two u16x2 SIMD pairs with the native
// VIMNMX3.U16x2 (3-input packed min/max): window-of-9 max = max3 of max3's, 40 packed ops per pair and ring
// polarity. Non-max suppression and the iniTh -> minTh retry run on a byte score map in shared memory, emission is
// ballot-ordered so the candidate order equals the serial reference. Integer-ALU bound, not HBM bound.
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

namespace orbx {

constexpr int kFastWarps = 4;
constexpr int kFastHead = 128;  // bytes in front of the per-warp regions: one mbarrier per warp

// Kernel parameters must stay below 4 KB for the descriptors to be usable from parameter space, so the TMA variant
// covers up to 8 levels — every configuration the reference ships (Examples/**/*.yaml: ORBextractor.nLevels = 8);
// more levels take the LDG variant.
constexpr int kFastMapLevels = 8;
struct FastMaps {
  CUtensorMap lv[kFastMapLevels];  // u8 [frames][h][w] view of every raw level; box = (box_w, box_h, 1)
};

struct FastSmemLayout {
  int tp;         // u16 row pitch of the raw tile (multiple of 4)
  int sp;         // byte row pitch of the score map (multiple of 4)
  int box_w;      // TMA box: bytes per row (multiple of 16) x rows; the u8 landing zone is reused as the score map
  int box_h;
  int raw_bytes;  // per warp
  int score_bytes;
  int list_bytes; // u16 per 4-pixel group: the groups that hold a pixel at or above the pass threshold
  int per_warp;   // multiple of 128 (TMA destination alignment)
};

// shared-memory layout that fits every cell of levels [l0, l1)
static FastSmemLayout fast_layout(const Plan& P, int l0, int l1) {
  int wc = 0, hc = 0;
  for (int l = l0; l < l1; l++) {
    if (P.lv[l].wCell > wc) wc = P.lv[l].wCell;
    if (P.lv[l].hCell > hc) hc = P.lv[l].hCell;
  }
  FastSmemLayout L;
  L.tp = round_up(wc + 12, 8);   // 16-byte rows: the widening pass stores 8 pixels at a time
  L.sp = round_up(wc + 8, 4);
  // 16-byte aligned start (up to 15 columns in front of the window) + the words the widening pass reads for the last
  // 8-pixel group that holds a window column
  const int need_px = wc + 6 + 15, need_words = 3 + 2 * ((wc + 5) / 8) + 3;
  L.box_w = round_up(need_px > 4 * need_words ? need_px : 4 * need_words, 16);
  L.box_h = hc + 6;
  L.raw_bytes = round_up(L.tp * (hc + 6) * 2, 128);
  L.score_bytes = round_up(L.sp * (hc + 2), 16);
  if (L.box_w * L.box_h > L.score_bytes) L.score_bytes = L.box_w * L.box_h;
  L.score_bytes = round_up(L.score_bytes, 128);
  L.list_bytes = round_up(2 * ((wc + 3) / 4) * hc, 128);
  L.per_warp = L.raw_bytes + L.score_bytes + L.list_bytes;
  return L;
}

size_t fast_smem_bytes(const Plan& P) { return kFastHead + (size_t)fast_layout(P, 0, P.nlevels).per_warp * kFastWarps; }

// ring pixel pair for the two pixels whose u16 columns are O, O+1 inside the 10-column window W (5 words)
template <int O, int N>
__device__ __forceinline__ uint32_t pair_at(const uint32_t (&W)[N]) {
  if constexpr ((O & 1) == 0) return W[O / 2];
  else return __byte_perm(W[(O - 1) / 2], W[(O + 1) / 2], 0x5432);
}

// Measured dead ends (tools/ubench/alu_tput.cu): every VIMNMX form, 2- or 3-input, issues at 64 lanes/clk/SM, so the
// window-of-3 / window-of-9 scheme below (80 three-input operations per pixel pair) beats a prefix / suffix (van
// Herk) split of the ring (57 two-input operations per polarity, measured 9 % slower in the kernel); HFMA2.RELU runs
// on the other pipe at the same rate but shares the issue port (a VIMNMX3 | HFMA2 mix reaches 82 lanes/clk/SM).
__device__ __forceinline__ void arc_minmax(const uint32_t (&v)[16], uint32_t& rhi, uint32_t& rlo) {
  uint32_t a[16], b[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    a[k] = __vimax3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
    b[k] = __vimin3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
  }
  uint32_t hi = 0xffffffffu, lo = 0u;
#pragma unroll
  for (int s = 0; s < 16; s += 2) {
    const uint32_t m0 = __vimax3_u16x2(a[s], a[(s + 3) & 15], a[(s + 6) & 15]);
    const uint32_t m1 = __vimax3_u16x2(a[s + 1], a[(s + 4) & 15], a[(s + 7) & 15]);
    hi = __vimin3_u16x2(hi, m0, m1);
    const uint32_t n0 = __vimin3_u16x2(b[s], b[(s + 3) & 15], b[(s + 6) & 15]);
    const uint32_t n1 = __vimin3_u16x2(b[s + 1], b[(s + 4) & 15], b[(s + 7) & 15]);
    lo = __vimax3_u16x2(lo, n0, n1);
  }
  rhi = hi;
  rlo = lo;
}

// FAST "m" of the two pixels packed in c (u16x2) -> r = max(m - tlow, 0) per u16 lane (0 <=> not a corner at tlow).
// All lanes hold values < 256, so a bias of 256 per lane keeps the packed subtractions borrow free.
__device__ __forceinline__ uint32_t score_pair(uint32_t c, uint32_t rhi, uint32_t rlo, uint32_t k_bias) {
  const uint32_t t1 = (c | 0x01000100u) - rhi;    // 256 + v - min_arcs max
  const uint32_t t2 = (rlo | 0x01000100u) - c;    // 256 + max_arcs min - v
  return __vmaxu2(__vmaxu2(t1, t2), k_bias) - k_bias;  // k_bias = 256 + tlow per lane
}

// 0x80 in every byte of w that is >= k (1 <= k <= 128; the caller handles larger k byte by byte): the low 7 bits
// carry into bit 7 exactly when they reach k, and a byte with bit 7 set is >= 128 >= k anyway.
__device__ __forceinline__ uint32_t bytes_ge(uint32_t w, int k) {
  if (k <= 128) return (((w & 0x7f7f7f7fu) + (uint32_t)(128 - k) * 0x01010101u) | w) & 0x80808080u;
  uint32_t m = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) m |= (uint32_t)((int)((w >> (8 * b)) & 0xff) >= k) << (8 * b + 7);
  return m;
}

// 3x3 non-max suppression of the 4 pixels of group g in interior row r: returns their score bytes where the pixel is
// a strict maximum of its 8 neighbours' raw scores, else 0. OpenCV keeps a corner at threshold T iff its score beats
// the scores of its neighbours that are corners at T; for a pixel that is itself a corner at T a neighbour below T can
// never beat it, so this one word serves every threshold. The bytes of the three score rows are split into u16x2
// lanes (even / odd pixels), neighbour maxima are VIMNMX3, the comparison is a biased subtraction.
__device__ __forceinline__ uint32_t nms_word(const uint8_t* score, int sp, int r, int g) {
  const uint32_t* mid = reinterpret_cast<const uint32_t*>(score + (r + 1) * sp + 4 + 4 * g);
  const int rw = sp >> 2;
  // per row: A = (p-1, p1), B = (p0, p2), Cc = (p1, p3), D = (p2, p4) as u16x2
  uint32_t A[3], B[3], Cc[3], D[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t* rowp = mid + (k - 1) * rw;
    const uint32_t wl = rowp[-1], wc = rowp[0], wr = rowp[1];
    A[k] = __byte_perm(wl, wc, 0x0503) & 0x00ff00ffu;
    B[k] = __byte_perm(wc, 0u, 0x4240);
    Cc[k] = __byte_perm(wc, 0u, 0x4341);
    D[k] = __byte_perm(wc, wr, 0x0402) & 0x00ff00ffu;
  }
  uint32_t ne = __vimax3_u16x2(__vimax3_u16x2(A[0], B[0], Cc[0]), __vimax3_u16x2(A[2], B[2], Cc[2]), A[1]);
  ne = __vimax3_u16x2(ne, Cc[1], Cc[1]);                       // neighbours of the even pixels (p0, p2)
  uint32_t no = __vimax3_u16x2(__vimax3_u16x2(B[0], Cc[0], D[0]), __vimax3_u16x2(B[2], Cc[2], D[2]), B[1]);
  no = __vimax3_u16x2(no, D[1], D[1]);                         // neighbours of the odd pixels (p1, p3)
  // lane = 0x8000 + neighbour max - self: bit 15 set <=> some neighbour >= self <=> not a strict local maximum
  const uint32_t te = ((ne | 0x80008000u) - B[1]) & 0x80008000u;
  const uint32_t to = ((no | 0x80008000u) - Cc[1]) & 0x80008000u;
  const uint32_t ke = B[1] & ~((te >> 15) * 0xffu);
  const uint32_t ko = Cc[1] & ~((to >> 15) * 0xffu);
  return __byte_perm(ke, ko, 0x6240);                          // bytes p0 p1 p2 p3
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// No minimum-blocks hint: with (128, 1) ptxas spends 111 registers instead of 80, 4 blocks per SM fit instead of the 6
// the 34 KB shared-memory layout allows, and the kernel is 10 % slower (3.90 -> 4.31 ms per 1024 pairs, measured).
template <bool kTma>
__global__ void __launch_bounds__(kFastWarps * 32)
k_fast(const __grid_constant__ Plan P, const __grid_constant__ FastMaps maps, const FrameSet fs, const WorkSet ws,
       int ini_th, int min_th, int tp, int sp, int raw_bytes, int score_bytes, int per_warp, int box_w,
       int box_bytes, int cell_begin, int cell_end, int pretest_mask) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = cell_begin + blockIdx.x * kFastWarps + warp;  // this launch covers the cells [cell_begin, cell_end)
  const int f = blockIdx.y;
  if (cell >= cell_end) return;
  int l = 0;
  while (l + 1 < P.nlevels && P.lv[l + 1].cell_base <= cell) l++;
  const LevelPlan& L = P.lv[l];
  const int ci = cell - L.cell_base;
  const int i = ci / L.nCols, j = ci - i * L.nCols;
  int32_t* count_out = ws.cell_count + (int64_t)f * P.cells_per_frame + cell;
  uint32_t* slot = ws.slots + (int64_t)f * P.slots_per_frame + L.slot_base + (int64_t)ci * L.slot_cap;

  // cell window (:909-921): [iniX, maxX) x [iniY, maxY) in level coordinates
  const int iniX = kMinBorder + j * L.wCell, iniY = kMinBorder + i * L.hCell;
  int maxX = iniX + L.wCell + 6, maxY = iniY + L.hCell + 6;
  if (maxX > L.maxBX) maxX = L.maxBX;
  if (maxY > L.maxBY) maxY = L.maxBY;
  const int tw = maxX - iniX, th = maxY - iniY;
  const int iw = tw - 6, ih = th - 6;  // pixels cv::FAST evaluates: a 3-px rim is skipped
  if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6 || iw <= 0 || ih <= 0) {
    if (lane == 0) *count_out = 0;
    return;
  }

  uint16_t* raw = reinterpret_cast<uint16_t*>(smem + kFastHead + (size_t)warp * per_warp);
  uint8_t* score = smem + kFastHead + (size_t)warp * per_warp + raw_bytes;
  uint16_t* list = reinterpret_cast<uint16_t*>(score + score_bytes);

  if constexpr (kTma) {
    // ---- one TMA tile copy per warp: box_w x box_h bytes of the level at (iniX & ~15, iniY, f) land in the (not yet
    //      used) score region; bytes outside the level read as zero. The box must START on a 16-byte boundary of the
    //      row (measured: an unaligned innermost coordinate raises an illegal-instruction fault, tools/ubench/
    //      tma_probe.cu), so the box is 15 columns wider than the window and the widening pass below starts at the
    //      byte offset iniX & 15. It only touches the th x tp window it needs — what lies to the right of the cell
    //      window are the neighbour cell's pixels, which only ever feed the scores of columns >= iw, and those are
    //      masked. ----
    const uint32_t bar = smem_u32(smem + 8 * warp), dst = smem_u32(score);
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&maps.lv[l])), "r"(iniX & ~15), "r"(iniY), "r"(f), "r"(bar)
          : "memory");
    }
    __syncwarp();
    uint32_t ok;
    do {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok) : "r"(bar) : "memory");
    } while (!ok);
    // widen u8 -> u16: a lane step cuts 8 pixels out of three landing-zone words (two PRMTs, the byte offset is a
    // per-warp constant) and turns them into one 16-byte store. tp is padded to a multiple of 8 for this.
    const int off = iniX & 15;
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(score) + (off >> 2);
    const uint32_t cut = 0x3210u + 0x1111u * (uint32_t)(off & 3);
    const int groups = tp >> 3, wpr = box_w >> 2, jmax = wpr - (off >> 2) - 2;  // words 2j .. 2j + 2 must exist
    const int q32 = 32 / groups, m32 = 32 - q32 * groups;
    int r = lane / groups, j = lane - r * groups;
    while (r < th) {
      uint32_t v0 = 0u, v1 = 0u;
      if (2 * j < jmax) {
        const uint32_t* p = t32 + r * wpr + 2 * j;
        const uint32_t a = p[0], b = p[1], c = p[2];
        v0 = __byte_perm(a, b, cut);
        v1 = __byte_perm(b, c, cut);
      }
      uint4 o;
      o.x = __byte_perm(v0, 0u, 0x4140);  // [b0, 0, b1, 0]
      o.y = __byte_perm(v0, 0u, 0x4342);  // [b2, 0, b3, 0]
      o.z = __byte_perm(v1, 0u, 0x4140);
      o.w = __byte_perm(v1, 0u, 0x4342);
      *reinterpret_cast<uint4*>(raw + r * tp + 8 * j) = o;
      j += m32;
      r += q32;
      if (j >= groups) {
        j -= groups;
        r++;
      }
    }
    __syncwarp();  // the landing zone becomes the score map
  } else {

  // ---- stage the raw window, widened to u16; columns >= tw are zero. A lane step produces 4 tile columns from the
  //      two aligned global words that hold them (funnel shift by the row's byte misalignment), one 8-byte store.
  //      The (row, group) walk is incremental (no division) and 4 lane steps are in flight at once: the loads of a
  //      batch are all issued before the first conversion, otherwise every step pays a full L2 / HBM round trip ----
  int pitch;
  const uint8_t* img = raw_level(P, fs, l, f, &pitch);
  const uint8_t* src = img + (int64_t)iniY * pitch + iniX;
  {
    const int groups = tp >> 2;
    const int q32 = 32 / groups, m32 = 32 - q32 * groups;
    int r = lane / groups, j = lane - r * groups;
    constexpr int kStageBatch = 4;
    while (r < th) {
      uint32_t lo[kStageBatch], hi[kStageBatch], sh[kStageBatch];
      int rr[kStageBatch], jj[kStageBatch], valid[kStageBatch];
#pragma unroll
      for (int u = 0; u < kStageBatch; u++) {
        rr[u] = r;
        jj[u] = j;
        valid[u] = r < th ? tw - 4 * j : 0;  // tile columns 4j .. 4j+3 that exist
        lo[u] = hi[u] = sh[u] = 0;
        if (valid[u] > 0) {
          const uint8_t* p = src + (int64_t)r * pitch + 4 * j;
          sh[u] = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
          const uint32_t* a = reinterpret_cast<const uint32_t*>(p - sh[u]);
          lo[u] = __ldg(a);
          if (sh[u]) hi[u] = __ldg(a + 1);  // stays inside the image row: the window ends >= 16 px before it
        }
        j += m32;
        r += q32;
        if (j >= groups) {
          j -= groups;
          r++;
        }
      }
#pragma unroll
      for (int u = 0; u < kStageBatch; u++) {
        if (rr[u] < th) {
          uint32_t v = __funnelshift_r(lo[u], hi[u], 8 * sh[u]);
          if (valid[u] <= 0) v = 0;
          else if (valid[u] < 4) v &= (1u << (8 * valid[u])) - 1u;
          uint2 o;
          o.x = __byte_perm(v, 0u, 0x4140);  // [b0, 0, b1, 0]
          o.y = __byte_perm(v, 0u, 0x4342);  // [b2, 0, b3, 0]
          *reinterpret_cast<uint2*>(raw + rr[u] * tp + 4 * jj[u]) = o;
        }
      }
    }
  }
  }
  // ---- clear the score map (1-row / 4-column zero frame around the interior) ----
  {
    uint32_t* s32 = reinterpret_cast<uint32_t*>(score);
    const int words = sp * (ih + 2) / 4;
    for (int k = lane; k < words; k += 32) s32[k] = 0u;
  }
  __syncwarp();

  // ---- scores, non-max suppression, per-cell threshold, ordered emission ----
  // The score map holds r = max(m - tlow, 0): a pixel is a corner at threshold T (m > T) iff r >= T - tlow + 1, its
  // OpenCV score is m - 1 = r + tlow - 1, and scores of 0 are never kept (cv::FAST compares with a strict >).
  //
  // Pass 0 (iniThFAST) does NOT score every pixel. A 9-arc of the 16-ring contains one end of every diameter, so a
  // corner at threshold T has max(p_k, p_k+8) > v + T on both compass diameters (or min < v - T on both): a 4-point
  // test that only ~2-7 % of the pixels of a natural image pass at T = 20. The groups of 4 pixels that contain such a
  // pixel are listed (ballot compaction keeps pixel order), only they get the 16-arc score, and of those only the
  // groups that reach the threshold go through non-max suppression and emission. Unlisted pixels keep score 0, which
  // is what a neighbour that is not a corner at T counts as in OpenCV's suppression. If the cell comes back empty
  // (:946) pass 1 scores every group at minThFAST the dense way.
  const int tlow = ini_th < min_th ? ini_th : min_th;
  const uint32_t k_bias = 0x01000100u + (uint32_t)tlow * 0x00010001u;
  const int gpr = (iw + 3) >> 2;
  const int ngroups = gpr * ih;
  const float inv_gpr = 1.0f / (float)gpr;
  const unsigned lt = (1u << lane) - 1u;
  const int rp = tp >> 1;  // words per tile row
  const int x_off = iniX + 3 - kMinBorder, y_off = iniY + 3 - kMinBorder;  // candidate coords are minBorder-relative
  int total = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int T = pass == 0 ? ini_th : min_th;
    const int r_min = max(max(T - tlow + 1, 2 - tlow), 1);  // m > T and m - 1 > 0, in the r domain
    int n_score = ngroups;
    const bool listed = (pretest_mask >> pass) & 1;
    if (listed) {
      // -- compass pre-test, two adjacent groups (8 pixels) per lane step: rows r, r+3, r+6 of the tile are
      //    dy = -3, 0, +3; the step's tile words come in with 16-byte loads (tile rows and 8-pixel starts are 16-byte
      //    aligned) --
      const uint32_t kb = 0x80008000u - (uint32_t)(min(max(T, 0), 255) + 1) * 0x00010001u;  // x >= y + T + 1 <=> bit 15 of x + kb - y
      n_score = 0;
      const int ppr = (gpr + 1) >> 1, npairs = ppr * ih;      // group pairs per row / per cell
      const int q32 = 32 / ppr, m32 = 32 - q32 * ppr;         // a step of 32 pairs = q32 rows + m32 pairs
      int r = lane / ppr, g2 = lane - r * ppr;
      for (int gbase = 0; gbase < npairs; gbase += 32) {
        bool hitA = false, hitB = false;
        if (gbase + lane < npairs) {
          const uint16_t* base = raw + r * tp + 8 * g2;
          uint32_t Wm3[6], W0[8], Wp3[6];
          {
            const uint4 a = *reinterpret_cast<const uint4*>(base), b = *reinterpret_cast<const uint4*>(base + 6 * tp);
            const uint2 a2 = *reinterpret_cast<const uint2*>(base + 8), b2 = *reinterpret_cast<const uint2*>(base + 6 * tp + 8);
            const uint4 c = *reinterpret_cast<const uint4*>(base + 3 * tp), c2 = *reinterpret_cast<const uint4*>(base + 3 * tp + 8);
            Wm3[0] = a.x; Wm3[1] = a.y; Wm3[2] = a.z; Wm3[3] = a.w; Wm3[4] = a2.x; Wm3[5] = a2.y;
            Wp3[0] = b.x; Wp3[1] = b.y; Wp3[2] = b.z; Wp3[3] = b.w; Wp3[4] = b2.x; Wp3[5] = b2.y;
            W0[0] = c.x; W0[1] = c.y; W0[2] = c.z; W0[3] = c.w; W0[4] = c2.x; W0[5] = c2.y; W0[6] = c2.z; W0[7] = c2.w;
          }
          auto test = [&](uint32_t p0, uint32_t p8, uint32_t p4, uint32_t p12, uint32_t c) {
            const uint32_t lo_of_hi = __vminu2(__vmaxu2(p0, p8), __vmaxu2(p4, p12));  // bright: must exceed v + T
            const uint32_t hi_of_lo = __vmaxu2(__vminu2(p0, p8), __vminu2(p4, p12));  // dark: must be below v - T
            // every lane stays in [0x8000 - 511, 0x8000 + 255], so bit 15 of x + 0x8000 - y says x >= y and nothing is
            // carried or borrowed across the two lanes; kb = 0x80008000 - K folds the bias and the threshold
            return (lo_of_hi + kb - c) | (c + kb - hi_of_lo);
          };
          const uint32_t bitsA = test(pair_at<3>(Wp3), pair_at<3>(Wm3), pair_at<6>(W0), pair_at<0>(W0), pair_at<3>(W0)) |
                                 test(pair_at<5>(Wp3), pair_at<5>(Wm3), pair_at<8>(W0), pair_at<2>(W0), pair_at<5>(W0));
          const uint32_t bitsB = test(pair_at<7>(Wp3), pair_at<7>(Wm3), pair_at<10>(W0), pair_at<4>(W0), pair_at<7>(W0)) |
                                 test(pair_at<9>(Wp3), pair_at<9>(Wm3), pair_at<12>(W0), pair_at<6>(W0), pair_at<9>(W0));
          hitA = (bitsA & 0x80008000u) != 0u;
          hitB = (bitsB & 0x80008000u) != 0u && 2 * g2 + 1 < gpr;
        }
        const unsigned balA = __ballot_sync(0xffffffffu, hitA), balB = __ballot_sync(0xffffffffu, hitB);
        const int pos = n_score + __popc(balA & lt) + __popc(balB & lt);
        if (hitA) list[pos] = (uint16_t)((r << 8) | (2 * g2));
        if (hitB) list[pos + (hitA ? 1 : 0)] = (uint16_t)((r << 8) | (2 * g2 + 1));
        n_score += __popc(balA) + __popc(balB);
        g2 += m32;
        r += q32;
        if (g2 >= ppr) {
          g2 -= ppr;
          r++;
        }
      }
      __syncwarp();
    }

    // -- 16-arc score of the listed groups (pass 0) or of every group (pass 1), 4 pixels per lane step; the groups
    //    that reach the pass threshold are listed again, in place (the write index never passes the read index) --
    int n_list = 0;
    for (int kbase = 0; kbase < n_score; kbase += 32) {
      const int k = kbase + lane;
      int r, g;
      if (listed) {
        const uint32_t e = k < n_score ? list[k] : 0u;
        r = (int)(e >> 8);
        g = (int)(e & 255u);
      } else {
        r = (int)(((float)k + 0.5f) * inv_gpr);
        g = k - r * gpr;
      }
      bool hit = false;
      if (k < n_score) {
        // rows r .. r+6 of the tile are dy = -3 .. 3 around interior row r; u16 columns 4g .. 4g+9
        const uint32_t* base = reinterpret_cast<const uint32_t*>(raw + r * tp + 4 * g);
        uint32_t Wm3[5], Wm2[5], Wm1[5], W0[5], Wp1[5], Wp2[5], Wp3[5];
        // a group starts on an 8-byte boundary of its tile row: two 8-byte loads + one word per row
        auto load_row = [&](int row, uint32_t (&W)[5]) {
          const uint2 a = *reinterpret_cast<const uint2*>(base + row * rp), b = *reinterpret_cast<const uint2*>(base + row * rp + 2);
          W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y;
          W[4] = base[row * rp + 4];
        };
        load_row(0, Wm3);
        load_row(1, Wm2);
        load_row(2, Wm1);
        load_row(3, W0);
        load_row(4, Wp1);
        load_row(5, Wp2);
        load_row(6, Wp3);
        uint32_t rr[2];
#pragma unroll
        for (int half = 0; half < 2; half++) {
          // ring k = 0..15: (dx, dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
          uint32_t v[16];
          uint32_t c;
          if (half == 0) {
            v[0] = pair_at<3>(Wp3);  v[1] = pair_at<4>(Wp3);  v[2] = pair_at<5>(Wp2);  v[3] = pair_at<6>(Wp1);
            v[4] = pair_at<6>(W0);   v[5] = pair_at<6>(Wm1);  v[6] = pair_at<5>(Wm2);  v[7] = pair_at<4>(Wm3);
            v[8] = pair_at<3>(Wm3);  v[9] = pair_at<2>(Wm3);  v[10] = pair_at<1>(Wm2); v[11] = pair_at<0>(Wm1);
            v[12] = pair_at<0>(W0);  v[13] = pair_at<0>(Wp1); v[14] = pair_at<1>(Wp2); v[15] = pair_at<2>(Wp3);
            c = pair_at<3>(W0);
          } else {
            v[0] = pair_at<5>(Wp3);  v[1] = pair_at<6>(Wp3);  v[2] = pair_at<7>(Wp2);  v[3] = pair_at<8>(Wp1);
            v[4] = pair_at<8>(W0);   v[5] = pair_at<8>(Wm1);  v[6] = pair_at<7>(Wm2);  v[7] = pair_at<6>(Wm3);
            v[8] = pair_at<5>(Wm3);  v[9] = pair_at<4>(Wm3);  v[10] = pair_at<3>(Wm2); v[11] = pair_at<2>(Wm1);
            v[12] = pair_at<2>(W0);  v[13] = pair_at<2>(Wp1); v[14] = pair_at<3>(Wp2); v[15] = pair_at<4>(Wp3);
            c = pair_at<5>(W0);
          }
          uint32_t rhi, rlo;
          arc_minmax(v, rhi, rlo);
          rr[half] = score_pair(c, rhi, rlo, k_bias);
        }
        uint32_t out = __byte_perm(rr[0], rr[1], 0x6420);  // low bytes of the 4 lanes = pixels 0..3
        // pixels beyond the interior width (last group of a row) must not score
        const int valid = iw - 4 * g;  // >= 1
        if (valid < 4) out &= (1u << (8 * valid)) - 1u;
        *reinterpret_cast<uint32_t*>(score + (r + 1) * sp + 4 + 4 * g) = out;
        hit = bytes_ge(out, r_min) != 0u;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      __syncwarp();
      if (hit) list[n_list + __popc(bal & lt)] = (uint16_t)((r << 8) | g);
      n_list += __popc(bal);
    }
    __syncwarp();

    // -- non-max suppression + threshold + emission on the listed groups. The list is in pixel row-major order, so
    //    the candidate order of the serial reference falls out of two ballots: no two kept pixels are adjacent, hence
    //    a group of 4 consecutive pixels keeps at most 2 --
    total = 0;
    for (int kbase = 0; kbase < n_list; kbase += 32) {
      const int k = kbase + lane;
      uint32_t word = 0u, keep = 0u;
      int r = 0, g = 0;
      if (k < n_list) {
        const uint32_t e = list[k];
        r = (int)(e >> 8);
        g = (int)(e & 255u);
        word = nms_word(score, sp, r, g);
        keep = bytes_ge(word, r_min);
      }
      const int c = __popc(keep);
      const unsigned b0 = __ballot_sync(0xffffffffu, c >= 1), b1 = __ballot_sync(0xffffffffu, c >= 2);
      if (keep) {
        int pos = total + __popc(b0 & lt) + __popc(b1 & lt);
#pragma unroll
        for (int q = 0; q < 4; q++)
          if ((keep >> (8 * q + 7)) & 1u)
            slot[pos++] = cand_pack(4 * g + q + x_off, r + y_off, (int)((word >> (8 * q)) & 0xff) + tlow - 1);
      }
      total += __popc(b0) + __popc(b1);
    }
    if (total > 0 || min_th == ini_th) break;  // :946 — retry with minThFAST only when the cell came back empty
    __syncwarp();  // the retry rewrites the score map and the list the suppression above has just read
  }
  if (lane == 0) *count_out = total;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Tensor maps of the raw levels of this batch. Returns false when a level cannot be described (base, pitch or frame
// stride not a 16-byte multiple — only possible for a caller-owned level 0): the LDG staging variant is used then.
static bool make_fast_maps(const Plan& P, const FrameSet& fs, const FastSmemLayout& L, int frames, FastMaps* M) {
  const EncodeTiledFn enc = encode_tiled();
  if (!enc || P.nlevels > kFastMapLevels) return false;
  for (int l = 0; l < P.nlevels; l++) {
    const uint8_t* base = l == 0 ? fs.lvl0 : fs.pyr + P.lv[l].img_off;
    const int64_t pitch = l == 0 ? fs.pitch0 : P.lv[l].pitch;
    int64_t fstride = l == 0 ? fs.fstride0 : fs.slab_fstride;
    if (frames == 1) fstride = (pitch * P.lv[l].h + 15) / 16 * 16;  // never applied
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (fstride & 15) || pitch <= 0 || fstride <= 0)
      return false;
    const cuuint64_t dims[3] = {(cuuint64_t)P.lv[l].w, (cuuint64_t)P.lv[l].h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)L.box_w, (cuuint32_t)L.box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&M->lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

// One launch for the levels [l0, l1) (their cells are consecutive in the per-frame numbering).
static void launch_fast_levels(const Plan& P, const FrameSet& fs, const WorkSet& ws, int ini_th, int min_th, int frames,
                               int l0, int l1, cudaStream_t st) {
  const FastSmemLayout L = fast_layout(P, l0, l1);
  const size_t smem = kFastHead + (size_t)L.per_warp * kFastWarps;
  const int attr = (int)(smem > 48 * 1024 ? smem : 48 * 1024);
  const int cell_begin = P.lv[l0].cell_base, cell_end = l1 < P.nlevels ? P.lv[l1].cell_base : P.cells_per_frame;
  dim3 grid((cell_end - cell_begin + kFastWarps - 1) / kFastWarps, frames);
  // bit p: pass p (0 = iniThFAST, 1 = the minThFAST retry) runs the compass pre-test and scores listed groups only.
  // Default: pass 0 only (DESIGN.md "measured decisions"); ORBX_FAST_PRETEST=3 is the A/B switch for the retry.
  static const int pretest_mask = getenv("ORBX_FAST_PRETEST") ? atoi(getenv("ORBX_FAST_PRETEST")) & 3 : 1;
  FastMaps M;
  if (L.box_w <= 256 && L.box_h <= 256 && make_fast_maps(P, fs, L, frames, &M)) {
    cudaFuncSetAttribute(k_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr);
    k_fast<true><<<grid, kFastWarps * 32, smem, st>>>(P, M, fs, ws, ini_th, min_th, L.tp, L.sp, L.raw_bytes,
                                                      L.score_bytes, L.per_warp, L.box_w, L.box_w * L.box_h,
                                                      cell_begin, cell_end, pretest_mask);
  } else {
    memset(&M, 0, sizeof(M));
    cudaFuncSetAttribute(k_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr);
    k_fast<false><<<grid, kFastWarps * 32, smem, st>>>(P, M, fs, ws, ini_th, min_th, L.tp, L.sp, L.raw_bytes,
                                                       L.score_bytes, L.per_warp, L.box_w, L.box_w * L.box_h,
                                                       cell_begin, cell_end, pretest_mask);
  }
}

// The small levels have a few tall cells (2 cell rows over ~100 pixel rows: hCell 51 where level 0 has 38) that would
// size the shared memory of EVERY warp. The levels are therefore cut into (at most) two launches at the point that
// minimises sum(cells / resident blocks per SM): at 752x480 levels 0-3 (86 % of the cells) run with 6 blocks per SM
// instead of 5.
static int fast_split_level(const Plan& P) {
  auto cost = [&](int l0, int l1) {
    const size_t smem = kFastHead + (size_t)fast_layout(P, l0, l1).per_warp * kFastWarps + 1024;  // + per-block reserve
    const int blocks = (int)((227u << 10) / smem);
    const int cells = (l1 < P.nlevels ? P.lv[l1].cell_base : P.cells_per_frame) - P.lv[l0].cell_base;
    return blocks > 0 ? (double)cells / blocks : 1e30;
  };
  int best_k = P.nlevels;
  double best = cost(0, P.nlevels);
  for (int k = 1; k < P.nlevels; k++) {
    const double c = cost(0, k) + cost(k, P.nlevels) + 4.0;  // a second launch must pay for itself
    if (c < best) {
      best = c;
      best_k = k;
    }
  }
  return best_k;
}

int fast_launch_count(const Plan& P) { return fast_split_level(P) < P.nlevels ? 2 : 1; }

void launch_fast(const Plan& P, const FrameSet& fs, const WorkSet& ws, int ini_th, int min_th, int frames,
                 cudaStream_t st) {
  const int k = fast_split_level(P);
  launch_fast_levels(P, fs, ws, ini_th, min_th, frames, 0, k, st);
  if (k < P.nlevels) launch_fast_levels(P, fs, ws, ini_th, min_th, frames, k, P.nlevels, st);
}

}  // namespace orbx
