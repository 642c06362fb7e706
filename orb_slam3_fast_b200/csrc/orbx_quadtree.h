// orbx_quadtree.h — ORBextractor::DistributeOctTree (src/ORBextractor.cc:557-757) + ExtractorNode::DivideNode
// (:490-540) + compareNodes (:542-555) as a data-parallel array algorithm executed by ONE WARP per (frame, level).
//
// Restructuring (not a translation of the std::list code):
//  * A node never stores its keypoints. Each candidate carries the list position of the node that owns it; splitting
//    a node is "every candidate of the node computes its quadrant" + a 4-bin histogram. DivideNode's stable partition
//    keeps candidate order inside every node, so "first key with the maximal response" (:739-754) is
//    argmax(response, -candidate index), which needs no ordering at all.
//  * The std::list is an array that is rebuilt once per round. In both phases of the reference the list after a
//    round is   reverse(children in creation order) ++ (old list without the split parents, order kept)
//    because children are push_front'ed (:621-672, :697-722) and parents erased in place. Positions are prefix sums
//    (ballot ranks / one blocked warp scan per round).
//  * Phase 2 (:678-735) needs the permutation libstdc++'s std::sort produces for nodes that tie on (size, UL.x).
//    std_sort_emulate_warp below reproduces it with the whole warp: every Hoare partition of the introsort is two
//    ordered compactions (the stop positions of the two scanning pointers) + one round of pairwise swaps, and the
//    final insertion sort is a stable rank inside a +-15 window. Lane 0 then walks from the back until the list holds
//    >= N nodes (:729).
//
//  * The per-CANDIDATE sweeps (root labels, first histogram, one sweep per round: 55 % of a frame's tree time, 70 % of a
//    level-0 tree, tools/qt_prof.py) are spread over the whole CTA ("team": the tree's warp 0 plus helper warps that
//    wait at a barrier for a command); everything per NODE — list rebuild, sort, walk — stays on warp 0. The sweeps only
//    touch their own candidate's label and shared-memory atomics, so the team needs no other coordination.
//
// The same source runs on the CPU (tests/hostcheck.cpp) with one "lane": ORBX_LANES loops cover every index, ballots
// degenerate to serial counters, the team is one thread. That is how the algorithm is checked against the oracle (and
// the sort against the real std::sort) without a GPU.
#ifndef ORBX_QUADTREE_H_
#define ORBX_QUADTREE_H_

#include "orbx_math.h"

namespace orbx {

#if defined(__CUDA_ARCH__)
#define ORBX_LANE() ((int)(threadIdx.x & 31))
#define ORBX_NLANES 32
#define ORBX_WSYNC() __syncwarp()
#define ORBX_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define ORBX_ATOMIC_MAX(p, v) atomicMax((p), (v))
#else
#define ORBX_LANE() 0
#define ORBX_NLANES 1
#define ORBX_WSYNC() (void)0
#define ORBX_ATOMIC_ADD(p, v) (*(p) += (v))
#define ORBX_ATOMIC_MAX(p, v) (*(p) = *(p) > (v) ? *(p) : (v))
#endif
#define ORBX_LANES(i, n) for (int i = ORBX_LANE(); i < (n); i += ORBX_NLANES)
// the team = every thread of the CTA (candidate sweeps); host: one thread
#if defined(__CUDA_ARCH__)
#define ORBX_TLANE() ((int)threadIdx.x)
#define ORBX_TNLANES ((int)blockDim.x)
#define ORBX_TSYNC() __syncthreads()
#else
#define ORBX_TLANE() 0
#define ORBX_TNLANES 1
#define ORBX_TSYNC() (void)0
#endif

// Shared-memory histogram / arg-max updates, predicated. (Warp-aggregating them with __match_any_sync was measured
// 15 % SLOWER on B200 than letting the hardware serialise same-address shared atomics.)
#define ORBX_AGG_ADD(base, key, valid)                \
  do {                                                \
    if (valid) ORBX_ATOMIC_ADD(&(base)[key], 1);      \
  } while (0)
#define ORBX_AGG_MAX(base, key, val, valid)                                  \
  do {                                                                       \
    if (valid) ORBX_ATOMIC_MAX((unsigned int*)&(base)[key], (unsigned)(val)); \
  } while (0)
// loops whose body contains a warp collective: the same trip count for every lane, `i` may run past n
#define ORBX_LANES_UNIFORM(i, n) for (int i = ORBX_LANE(); i - ORBX_LANE() < (n); i += ORBX_NLANES)

// candidate word: x:12 | y:12 | score:8 (x, y relative to minBorder, as the reference's vToDistributeKeys)
ORBX_HD uint32_t cand_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
ORBX_HD int cand_x(uint32_t c) { return (int)(c & 0xfff); }
ORBX_HD int cand_y(uint32_t c) { return (int)((c >> 12) & 0xfff); }
ORBX_HD int cand_s(uint32_t c) { return (int)(c >> 24); }

// Optional cycle accounting (build with -DORBX_QT_PROF): lane 0 adds the cycles since the previous mark to a category.
#if defined(ORBX_QT_PROF) && defined(__CUDA_ARCH__)
#define ORBX_QT_MARK(T, cat)                                \
  do {                                                      \
    if (ORBX_LANE() == 0 && (T).prof) {                     \
      const long long now_ = clock64();                     \
      (T).prof[cat] += now_ - (T).prof_prev;                \
      (T).prof_prev = now_;                                 \
    }                                                       \
  } while (0)
#else
#define ORBX_QT_MARK(T, cat) (void)0
#endif
enum { kQtInit = 2, kQtSelect = 3, kQtSort = 4, kQtWalk = 5, kQtChildren = 6, kQtPrep = 7, kQtSweep = 8, kQtRounds = 10,
       kQtRounds2 = 11, kQtProfSlots = 16 };

// ---- warp primitives with a serial twin ---------------------------------------------------------------------------
// Indices i in [0, n) with pred(i), ascending, are written to out[0..count); returns count. If rank_of is not null,
// rank_of[i] receives the rank of every selected i (others untouched).
template <class Pred>
ORBX_HD int warp_compact(int n, uint16_t* out, Pred pred) {
#if defined(__CUDA_ARCH__)
  const int lane = ORBX_LANE();
  const unsigned lt = (1u << lane) - 1u;
  int base = 0;
  for (int b = 0; b < n; b += 32) {
    const int i = b + lane;
    const bool f = i < n && pred(i);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (f) out[base + __popc(m & lt)] = (uint16_t)i;
    base += __popc(m);
  }
  __syncwarp();
  return base;
#else
  int c = 0;
  for (int i = 0; i < n; i++)
    if (pred(i)) out[c++] = (uint16_t)i;
  return c;
#endif
}

template <class Pred>
ORBX_HD int warp_count(int n, Pred pred) {
#if defined(__CUDA_ARCH__)
  int c = 0;
  for (int b = 0; b < n; b += 32) {
    const int i = b + ORBX_LANE();
    c += __popc(__ballot_sync(0xffffffffu, i < n && pred(i)));
  }
  return c;
#else
  int c = 0;
  for (int i = 0; i < n; i++) c += pred(i) ? 1 : 0;
  return c;
#endif
}

// exclusive prefix sum of a[0..n) in place, returns the total. Device: every lane owns a contiguous block of `per`
// elements (per odd: conflict-free shared-memory strides), one shuffle scan over the 32 block sums.
ORBX_HD int excl_scan(int* a, int n) {
#if defined(__CUDA_ARCH__)
  const int lane = ORBX_LANE();
  const int per = ((n + 31) >> 5) | 1;
  const int b0 = lane * per;
  int sum = 0;
  for (int k = 0; k < per; k++) sum += (b0 + k < n) ? a[b0 + k] : 0;
  int inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  const int total = __shfl_sync(0xffffffffu, inc, 31);
  int run = inc - sum;
  for (int k = 0; k < per; k++) {
    if (b0 + k < n) {
      const int v = a[b0 + k];
      a[b0 + k] = run;
      run += v;
    }
  }
  __syncwarp();
  return total;
#else
  int s = 0;
  for (int i = 0; i < n; i++) {
    const int v = a[i];
    a[i] = s;
    s += v;
  }
  return s;
#endif
}

// ---- libstdc++ std::sort, warp-cooperative ------------------------------------------------------------------------
// Produces exactly the permutation of std_sort_emulate (orbx_math.h), i.e. of std::sort in bits/stl_algo.h.
//  * __unguarded_partition(first + 1, last, pivot = first): the left pointer stops at the elements with !(a < pivot),
//    the right pointer at the elements with !(pivot < a). Until the pointers cross, both only look at positions that
//    no swap has touched yet, so the k-th left stop i_k / right stop j_k are the k-th such positions of the ORIGINAL
//    segment counted from the left / right. Pairs are swapped while i_k < j_k (K pairs, a prefix because i grows and
//    j shrinks). The returned cut is the left pointer's next stop: the next original stop i_K or the position j_{K-1}
//    that has just received an element >= pivot, whichever comes first.
//  * The ranges left by __introsort_loop are <= 16 long (or heap-sorted), and every element of a range is <= every
//    element of the ranges to its right; __final_insertion_sort is a stable insertion sort, so the final index of
//    element i is  #{j : key_j < key_i  or  (key_j == key_i and j < i)}  and only j in [i - 15, i + 15] can differ
//    from "j < i".
struct SortScratch {
  uint16_t* lidx;  // [n]
  uint16_t* ridx;  // [n]
  SortElem* tmp;   // [n]
};

ORBX_HD int ss_partition_warp(SortElem* a, int lo, int hi, const SortScratch& W) {
  const uint32_t p = a[lo].key;
  const int m = hi - lo - 1;  // partitioned range: lo + 1 .. hi - 1
  const int nL = warp_compact(m, W.lidx, [&](int t) { return !(a[lo + 1 + t].key < p); });
  const int nR = warp_compact(m, W.ridx, [&](int t) { return !(p < a[hi - 1 - t].key); });
  const int nmin = nL < nR ? nL : nR;
  const int K = warp_count(nmin, [&](int k) { return lo + 1 + (int)W.lidx[k] < hi - 1 - (int)W.ridx[k]; });
  int cut = hi;
  if (K < nL) cut = lo + 1 + (int)W.lidx[K];
  if (K >= 1) {
    const int j = hi - 1 - (int)W.ridx[K - 1];
    if (j < cut) cut = j;
  }
  ORBX_WSYNC();
  ORBX_LANES(k, K) se_swap(a[lo + 1 + (int)W.lidx[k]], a[hi - 1 - (int)W.ridx[k]]);
  ORBX_WSYNC();
  return cut;
}

ORBX_HD void ss_move_median_to_first_warp(SortElem* a, int result, int ia, int ib, int ic) {
  const uint32_t ka = a[ia].key, kb = a[ib].key, kc = a[ic].key;
  int pick;
  if (ka < kb) pick = kb < kc ? ib : (ka < kc ? ic : ia);
  else pick = ka < kc ? ia : (kb < kc ? ic : ib);
  ORBX_WSYNC();
  if (ORBX_LANE() == 0) se_swap(a[result], a[pick]);
  ORBX_WSYNC();
}

ORBX_HD void std_sort_emulate_warp(SortElem* a, int n, const SortScratch& W) {
  if (n <= 1) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) lg++;
  int stack[3 * 48];  // (first, last, depth); at most one entry per partition level of the current path
  int sp = 0;
  stack[sp++] = 0;
  stack[sp++] = n;
  stack[sp++] = 2 * lg;
  while (sp > 0) {
    int depth = stack[--sp];
    int last = stack[--sp];
    int first = stack[--sp];
    while (last - first > 16) {
      if (depth == 0) {
        ORBX_WSYNC();
        if (ORBX_LANE() == 0) {
          ORBX_SORT_HEAP_HOOK;
          ss_heap_sort(a + first, last - first);
        }
        ORBX_WSYNC();
        break;
      }
      --depth;
      ss_move_median_to_first_warp(a, first, first + 1, first + (last - first) / 2, last - 1);
      const int cut = ss_partition_warp(a, first, last, W);
      // the original recurses into [cut, last) and loops on [first, cut); the ranges are disjoint, the order of
      // processing does not change the result. Push the larger one so the stack stays logarithmic.
      if (last - cut > cut - first) {
        stack[sp++] = cut; stack[sp++] = last; stack[sp++] = depth;
        last = cut;
      } else {
        stack[sp++] = first; stack[sp++] = cut; stack[sp++] = depth;
        first = cut;
      }
    }
  }
  ORBX_WSYNC();
  // __final_insertion_sort == stable sort of ranges that are already ordered among themselves
  ORBX_LANES(i, n) {
    const SortElem e = a[i];
    const int jlo = i - 15 > 0 ? i - 15 : 0, jhi = i + 15 < n - 1 ? i + 15 : n - 1;
    int r = jlo;
    for (int j = jlo; j <= jhi; j++) {
      const uint32_t kj = a[j].key;
      r += (kj < e.key || (kj == e.key && j < i)) ? 1 : 0;
    }
    W.tmp[r] = e;
  }
  ORBX_WSYNC();
  ORBX_LANES(i, n) a[i] = W.tmp[i];
  ORBX_WSYNC();
}

// ---- the tree -----------------------------------------------------------------------------------------------------
struct QBox {
  int16_t ulx, urx, uly, bry;
};

// Working set of one tree. All arrays have `cap` entries unless noted; they live in shared memory on the device.
struct QTree {
  int cap;
  QBox* box[2];        // list (double buffered)
  int* cnt[2];         // keys per node
  int* child[2];       // [cap * 4] histogram of the splittable nodes' candidates over the 4 quadrants
  uint16_t* newpos;    // old position -> new position (survivors)
  uint16_t* childpos;  // [cap * 4] old position, quadrant -> new position (split parents); sort scratch in between
  uint8_t* committed;  // old position was split this round
  uint8_t* splittable; // [2][cap] node takes part in the next round's histogram
  uint16_t* pending[2];// positions of the nodes the next phase-2 round may split, in creation order
  SortElem* sortbuf;
  uint16_t* rank2pos;  // creation rank -> old position of the parent, [cap]
  int* scan;           // [cap + 1] scratch for prefix sums
  int* vars;           // [8] warp-uniform scalars written by lane 0
  // candidates: shared memory when they fit, else global memory
  const uint32_t* cand;
  uint16_t* lab;       // per candidate: node position | quadrant << 14
  int C;
  long long* prof;     // [kQtProfSlots] or null (ORBX_QT_PROF builds)
  long long prof_prev;
};

constexpr int kLabPosMask = 0x3fff;  // cap < 16384

ORBX_HD int quadrant_of(uint32_t c, const QBox& b) {
  const int mx = b.ulx + ((b.urx - b.ulx + 1) >> 1);  // UL.x + ceil((UR.x - UL.x) / 2)   :494
  const int my = b.uly + ((b.bry - b.uly + 1) >> 1);  // UL.y + ceil((BR.y - UL.y) / 2)   :495
  return (cand_x(c) < mx ? 0 : 1) + (cand_y(c) < my ? 0 : 2);  // n1, n2, n3, n4       :525-533
}

ORBX_HD QBox child_box(const QBox& b, int q) {
  const int mx = b.ulx + ((b.urx - b.ulx + 1) >> 1);
  const int my = b.uly + ((b.bry - b.uly + 1) >> 1);
  QBox c;
  c.ulx = (int16_t)((q & 1) ? mx : b.ulx);
  c.urx = (int16_t)((q & 1) ? b.urx : mx);
  c.uly = (int16_t)((q & 2) ? my : b.uly);
  c.bry = (int16_t)((q & 2) ? b.bry : my);
  return c;
}

ORBX_HD int nonempty4(const int* h) { return (h[0] > 0) + (h[1] > 0) + (h[2] > 0) + (h[3] > 0); }
ORBX_HD int multi4(const int* h) { return (h[0] > 1) + (h[1] > 1) + (h[2] > 1) + (h[3] > 1); }

// ---- the candidate sweeps, executed by the whole team ----------------------------------------------------------
enum { kQtCmdExit = 0, kQtCmdRoots = 1, kQtCmdFirst = 2, kQtCmdRound = 3 };
enum { kQtVarCmd = 4, kQtVarNxt = 5, kQtVarFinish = 6, kQtVarTotal = 7 };  // T.vars slots of the team protocol

// root of every candidate (:588-594) + root counts
ORBX_HD void qt_sweep_roots(const QTree& T, float hX) {
  for (int c = ORBX_TLANE(); c < T.C; c += ORBX_TNLANES) {
    const int r = (int)fdiv((float)cand_x(T.cand[c]), hX);
    T.lab[c] = (uint16_t)r;
    ORBX_ATOMIC_ADD(&T.cnt[1][r], 1);
  }
}

// compacted root position + quadrant inside it, first histogram
ORBX_HD void qt_sweep_first(const QTree& T) {
  for (int c = ORBX_TLANE(); c < T.C; c += ORBX_TNLANES) {
    const int p = T.newpos[T.lab[c]];
    uint32_t l = (uint32_t)p;
    if (T.splittable[p]) {
      const int q = quadrant_of(T.cand[c], T.box[0][p]);
      l |= (uint32_t)q << 14;
      ORBX_ATOMIC_ADD(&T.child[0][4 * p + q], 1);
    }
    T.lab[c] = (uint16_t)l;
  }
}

// one round: move every candidate to its new node, then histogram (or, in the final round, arg-max). Stage-wise over 4
// team steps so that the dependent shared-memory loads of different steps overlap:
// label -> new position (one table for split parents and survivors) -> node word -> quadrant -> histogram.
ORBX_HD void qt_sweep_round(const QTree& T, int nxt, bool finish) {
  const int C = T.C;
  int* child_nxt = T.child[nxt];
  const int step = ORBX_TNLANES, me = ORBX_TLANE();
  for (int c0 = 0; c0 < C; c0 += 4 * step) {
    uint32_t lv[4], cv[4];
    int np[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int c = c0 + u * step + me;
      ok[u] = c < C;
      lv[u] = ok[u] ? (uint32_t)T.lab[c] : 0u;
      cv[u] = ok[u] ? T.cand[c] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) np[u] = T.childpos[4 * (int)(lv[u] & kLabPosMask) + (int)(lv[u] >> 14)];
    if (finish) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = c0 + u * step + me;
        const uint32_t v = ((uint32_t)cand_s(cv[u]) << 24) | (0xffffffu - (uint32_t)c);
        ORBX_AGG_MAX(child_nxt, np[u], v, ok[u]);
        if (ok[u]) T.lab[c] = (uint16_t)np[u];
      }
    } else {
      uint32_t info[4];
#pragma unroll
      for (int u = 0; u < 4; u++) info[u] = (uint32_t)T.scan[np[u]];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = c0 + u * step + me;
        const bool split = ok[u] && (info[u] >> 31);
        const int q = (cand_x(cv[u]) < (int)(info[u] & 0xfff) ? 0 : 1) + (cand_y(cv[u]) < (int)((info[u] >> 12) & 0xfff) ? 0 : 2);
        ORBX_AGG_ADD(child_nxt, 4 * np[u] + q, split);
        if (ok[u]) T.lab[c] = (uint16_t)((uint32_t)np[u] | (split ? (uint32_t)q << 14 : 0u));
      }
    }
  }
}

ORBX_HD void qt_dispatch(const QTree& T, int cmd, int nxt, bool finish, float hX) {
  if (cmd == kQtCmdRoots) qt_sweep_roots(T, hX);
  else if (cmd == kQtCmdFirst) qt_sweep_first(T);
  else if (cmd == kQtCmdRound) qt_sweep_round(T, nxt, finish);
}

// warp 0: publish the command, run it with the team, return when every member is done
ORBX_HD void qt_team_run(const QTree& T, int cmd, int nxt, bool finish, float hX) {
  if (ORBX_LANE() == 0) {
    T.vars[kQtVarCmd] = cmd;
    T.vars[kQtVarNxt] = nxt;
    T.vars[kQtVarFinish] = finish ? 1 : 0;
  }
  ORBX_TSYNC();
  qt_dispatch(T, cmd, nxt, finish, hX);
  ORBX_TSYNC();
}

#if defined(__CUDACC__)
// the helper warps of a tree: wait for a command, run their share of the sweep, until warp 0 says exit
__device__ __forceinline__ void quadtree_helper(const QTree& T, float hX) {
  for (;;) {
    __syncthreads();
    const int cmd = T.vars[kQtVarCmd], nxt = T.vars[kQtVarNxt], fin = T.vars[kQtVarFinish];
    if (cmd == kQtCmdExit) return;
    qt_dispatch(T, cmd, nxt, fin != 0, hX);
    __syncthreads();
  }
}
// warp 0, after quadtree_run: release the helpers
__device__ __forceinline__ void quadtree_release(const QTree& T) {
  __syncwarp();
  if (ORBX_LANE() == 0) T.vars[kQtVarCmd] = kQtCmdExit;
  __syncthreads();
}
#endif

// Runs the whole culling for one level. Returns the number of selected keypoints; out[i] = candidate index of the
// i-th keypoint in list order (front to back). `width`, `height` = maxBorder - minBorder of the level.
ORBX_HD int quadtree_run(QTree& T, int width, int height, int nIni, float hX, int N, uint32_t* out_cand_idx) {
  const int C = T.C;
  if (C == 0) return 0;
  int cur = 0;
  // ---- roots (:566-601) ----
  ORBX_LANES(i, nIni) {
    QBox b;
    b.ulx = (int16_t)(int)fmul(hX, (float)i);
    b.urx = (int16_t)(int)fmul(hX, (float)(i + 1));
    b.uly = 0;
    b.bry = (int16_t)height;
    T.box[1][i] = b;  // staged in the other buffer, compacted below
    T.cnt[1][i] = 0;
  }
  ORBX_WSYNC();
  qt_team_run(T, kQtCmdRoots, 0, false, hX);
  int S = warp_compact(nIni, T.rank2pos, [&](int i) { return T.cnt[1][i] > 0; });
  ORBX_LANES(p, S) {
    const int i = T.rank2pos[p];
    T.box[0][p] = T.box[1][i];
    T.cnt[0][p] = T.cnt[1][i];
    T.newpos[i] = (uint16_t)p;
    T.splittable[p] = T.cnt[1][i] > 1;
    T.child[0][4 * p + 0] = 0;
    T.child[0][4 * p + 1] = 0;
    T.child[0][4 * p + 2] = 0;
    T.child[0][4 * p + 3] = 0;
  }
  ORBX_WSYNC();
  qt_team_run(T, kQtCmdFirst, 0, false, hX);
  ORBX_QT_MARK(T, kQtInit);

  bool phase2 = false;
  int n_pending = 0;  // entries of T.pending[cur]
  bool finish = false;
  while (!finish) {
    const int nxt = cur ^ 1;
    const QBox* box = T.box[cur];
    const int* cnt = T.cnt[cur];
    const int* child = T.child[cur];
    const uint8_t* split_cur = T.splittable + cur * T.cap;
    uint8_t* split_nxt = T.splittable + nxt * T.cap;
    const int prevS = S;
    int n_parents = 0;  // parents split this round; T.rank2pos[r] = position of the r-th created parent
    // ---- choose the parents and their creation order ----
    if (!phase2) {
      // phase 1 (:610-672): every non-leaf node, in list order
      n_parents = warp_compact(S, T.rank2pos, [&](int p) { return split_cur[p] != 0; });
      ORBX_LANES(p, S) T.committed[p] = split_cur[p];
      ORBX_WSYNC();
      ORBX_QT_MARK(T, kQtSelect);
    } else {
      // phase 2 (:679-735): sort the pending nodes by (size, UL.x) with std::sort, walk from the back, stop at N
      const uint16_t* pend = T.pending[cur];
      ORBX_LANES(p, S) T.committed[p] = 0;
      ORBX_LANES(i, n_pending) {
        const int p = pend[i];
        SortElem e;
        e.key = ((uint32_t)cnt[p] << 12) | (uint32_t)(uint16_t)box[p].ulx;
        e.id = (uint32_t)p;
        T.sortbuf[i] = e;
      }
      ORBX_WSYNC();
      SortScratch W;
      W.lidx = T.childpos;
      W.ridx = T.childpos + T.cap;
      W.tmp = reinterpret_cast<SortElem*>(T.box[nxt]);  // sizeof(SortElem) == sizeof(QBox); not live until below
      std_sort_emulate_warp(T.sortbuf, n_pending, W);
      ORBX_QT_MARK(T, kQtSort);
      if (ORBX_LANE() == 0) {
        int size = S, t = 0;
        for (int j = n_pending - 1; j >= 0; j--) {
          const int p = (int)T.sortbuf[j].id;
          T.committed[p] = 1;
          T.rank2pos[t++] = (uint16_t)p;
          size += nonempty4(child + 4 * p) - 1;
          if (size >= N) break;
        }
        T.vars[0] = t;
      }
      ORBX_WSYNC();
      n_parents = T.vars[0];
      ORBX_WSYNC();
      ORBX_QT_MARK(T, kQtWalk);
#if defined(ORBX_QT_PROF) && defined(__CUDA_ARCH__)
      if (ORBX_LANE() == 0 && T.prof) T.prof[kQtRounds2] += 1;
#endif
    }
    // ---- children: creation rank r -> first child index (low half) and first pending index (high half) ----
    ORBX_LANES(r, n_parents) {
      const int* h = child + 4 * (int)T.rank2pos[r];
      T.scan[r] = nonempty4(h) | (multi4(h) << 16);
    }
    ORBX_WSYNC();
    const int tot = excl_scan(T.scan, n_parents);
    const int K = tot & 0xffff, n_expand = tot >> 16;
    // children boxes / counts at their final position (list = reverse(children) ++ survivors); the ones with more than
    // one key form the pending list of the next phase-2 round, in creation order (:636-665, :700-720)
    int* child_nxt = T.child[nxt];
    ORBX_LANES(r, n_parents) {
      const int p = T.rank2pos[r];
      int k = T.scan[r] & 0xffff, e = T.scan[r] >> 16;
      const QBox pb = box[p];
      for (int q = 0; q < 4; q++) {
        const int n = child[4 * p + q];
        if (n > 0) {
          const int pos = K - 1 - k;
          T.box[nxt][pos] = child_box(pb, q);
          T.cnt[nxt][pos] = n;
          T.childpos[4 * p + q] = (uint16_t)pos;
          if (n > 1) T.pending[nxt][e++] = (uint16_t)pos;
          k++;
        }
      }
    }
    // survivors keep their relative order behind the children
#if defined(__CUDA_ARCH__)
    {
      const int lane = ORBX_LANE();
      const unsigned lt = (1u << lane) - 1u;
      int base = K;
      for (int b = 0; b < S; b += 32) {
        const int p = b + lane;
        const bool f = p < S && !T.committed[p];
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) {
          const int pos = base + __popc(m & lt);
          T.box[nxt][pos] = box[p];
          T.cnt[nxt][pos] = cnt[p];
          T.childpos[4 * p + 0] = T.childpos[4 * p + 1] = T.childpos[4 * p + 2] = T.childpos[4 * p + 3] = (uint16_t)pos;
        }
        base += __popc(m);
      }
      S = base;
    }
#else
    {
      int base = K;
      for (int p = 0; p < S; p++) {
        if (!T.committed[p]) {
          T.box[nxt][base] = box[p];
          T.cnt[nxt][base] = cnt[p];
          T.childpos[4 * p + 0] = T.childpos[4 * p + 1] = T.childpos[4 * p + 2] = T.childpos[4 * p + 3] = (uint16_t)base;
          base++;
        }
      }
      S = base;
    }
#endif
    ORBX_WSYNC();
    ORBX_QT_MARK(T, kQtChildren);
    // ---- termination (:676-678, :731-733) ----
    bool next_phase2 = phase2;
    if (S >= N || S == prevS) finish = true;
    else if (!phase2 && S + 3 * n_expand > N) next_phase2 = true;
    // ---- who is splittable in the next round ----
    if (!finish) {
      if (next_phase2) {
        ORBX_LANES(p, S) split_nxt[p] = 0;
        ORBX_WSYNC();
        ORBX_LANES(i, n_expand) split_nxt[T.pending[nxt][i]] = 1;
      } else {
        ORBX_LANES(p, S) split_nxt[p] = T.cnt[nxt][p] > 1;
      }
      ORBX_WSYNC();
      // one word per node for the sweep: split point of the nodes that take part in the next histogram
      ORBX_LANES(p, S) {
        child_nxt[4 * p + 0] = 0;
        child_nxt[4 * p + 1] = 0;
        child_nxt[4 * p + 2] = 0;
        child_nxt[4 * p + 3] = 0;
        const QBox b = T.box[nxt][p];
        const uint32_t mx = (uint32_t)(b.ulx + ((b.urx - b.ulx + 1) >> 1)), my = (uint32_t)(b.uly + ((b.bry - b.uly + 1) >> 1));
        T.scan[p] = split_nxt[p] ? (int)(0x80000000u | mx | (my << 12)) : 0;
      }
    } else {
      // final round: child_nxt[p] doubles as "best candidate" accumulator (response << 24 | ~index)
      ORBX_LANES(p, S) child_nxt[p] = 0;
    }
    ORBX_WSYNC();
    ORBX_QT_MARK(T, kQtPrep);
    // ---- one sweep over the candidates (the whole team): move to the new node, then histogram / argmax ----
    qt_team_run(T, kQtCmdRound, nxt, finish, hX);
    ORBX_WSYNC();
    ORBX_QT_MARK(T, kQtSweep);
#if defined(ORBX_QT_PROF) && defined(__CUDA_ARCH__)
    if (ORBX_LANE() == 0 && T.prof) T.prof[kQtRounds] += 1;
#endif
    phase2 = next_phase2;
    n_pending = n_expand;
    cur = nxt;
  }
  // ---- retain the best point of every node, list order (:739-754) ----
  const unsigned int* best = (const unsigned int*)T.child[cur];
  ORBX_LANES(p, S) out_cand_idx[p] = 0xffffffu - (best[p] & 0xffffffu);
  ORBX_WSYNC();
  return S;
}

}  // namespace orbx

#endif  // ORBX_QUADTREE_H_
