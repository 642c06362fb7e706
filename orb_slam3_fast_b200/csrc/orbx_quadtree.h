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
//    because children are push_front'ed (:621-672, :697-722) and parents erased in place. Positions are prefix sums.
//  * Phase 2 (:678-735) needs the permutation libstdc++'s std::sort produces for nodes that tie on (size, UL.x); lane 0
//    runs the emulation from orbx_math.h, then walks from the back until the list holds >= N nodes (:729).
//
// The same source runs on the CPU (tests/hostcheck.cpp) with one "lane": ORBX_LANES loops cover every index, scans are
// serial. That is how the algorithm is checked against the oracle without a GPU.
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

// candidate word: x:12 | y:12 | score:8 (x, y relative to minBorder, as the reference's vToDistributeKeys)
ORBX_HD uint32_t cand_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
ORBX_HD int cand_x(uint32_t c) { return (int)(c & 0xfff); }
ORBX_HD int cand_y(uint32_t c) { return (int)((c >> 12) & 0xfff); }
ORBX_HD int cand_s(uint32_t c) { return (int)(c >> 24); }

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
  uint16_t* childpos;  // [cap * 4] old position, quadrant -> new position (split parents)
  uint8_t* committed;  // old position was split this round
  uint8_t* splittable; // [2][cap] node takes part in the next round's histogram
  uint16_t* pending[2];// positions of the nodes the next phase-2 round may split, in creation order
  SortElem* sortbuf;
  uint16_t* rank2pos;  // creation rank -> old position of the parent (phase 2), [cap]
  int* scan;           // [cap + 1] scratch for prefix sums
  int* vars;           // [8] warp-uniform scalars written by lane 0
  // candidates (global memory on the device)
  const uint32_t* cand;
  uint32_t* lab;       // per candidate: node position | quadrant << 16
  int C;
};

// exclusive prefix sum of a[0..n) in place, returns the total (warp-cooperative on the device)
ORBX_HD int excl_scan(int* a, int n) {
#if defined(__CUDA_ARCH__)
  const int lane = ORBX_LANE();
  int carry = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const int v = i < n ? a[i] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (i < n) a[i] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  return carry;
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
  ORBX_LANES(c, C) {
    const int r = (int)fdiv((float)cand_x(T.cand[c]), hX);
    T.lab[c] = (uint32_t)r;
    ORBX_ATOMIC_ADD(&T.cnt[1][r], 1);
  }
  ORBX_WSYNC();
  ORBX_LANES(i, nIni) T.scan[i] = T.cnt[1][i] > 0 ? 1 : 0;
  ORBX_WSYNC();
  int S = excl_scan(T.scan, nIni);
  ORBX_LANES(i, nIni) {
    if (T.cnt[1][i] > 0) {
      const int p = T.scan[i];
      T.box[0][p] = T.box[1][i];
      T.cnt[0][p] = T.cnt[1][i];
      T.newpos[i] = (uint16_t)p;
    }
  }
  ORBX_WSYNC();
  ORBX_LANES(i, S) {
    T.splittable[i] = T.cnt[0][i] > 1;
    T.child[0][4 * i + 0] = 0;
    T.child[0][4 * i + 1] = 0;
    T.child[0][4 * i + 2] = 0;
    T.child[0][4 * i + 3] = 0;
  }
  ORBX_WSYNC();
  ORBX_LANES(c, C) {
    const int p = T.newpos[T.lab[c]];
    uint32_t l = (uint32_t)p;
    if (T.splittable[p]) {
      const int q = quadrant_of(T.cand[c], T.box[0][p]);
      l |= (uint32_t)q << 16;
      ORBX_ATOMIC_ADD(&T.child[0][4 * p + q], 1);
    }
    T.lab[c] = l;
  }
  ORBX_WSYNC();

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
    int K = 0;         // children created this round
    int n_parents = 0; // parents split this round
    // ---- choose the parents and their creation order ----
    if (!phase2) {
      // phase 1 (:610-672): every non-leaf node, in list order
      ORBX_LANES(p, S) {
        T.committed[p] = split_cur[p];
        T.scan[p] = split_cur[p] ? 1 : 0;
      }
      ORBX_WSYNC();
      n_parents = excl_scan(T.scan, S);
      ORBX_LANES(p, S) if (T.committed[p]) T.rank2pos[T.scan[p]] = (uint16_t)p;
      ORBX_WSYNC();
    } else {
      // phase 2 (:679-735): sort the pending nodes by (size, UL.x) with std::sort, walk from the back, stop at N
      ORBX_LANES(p, S) T.committed[p] = 0;
      ORBX_WSYNC();
      if (ORBX_LANE() == 0) {
        const uint16_t* pend = T.pending[cur];
        for (int i = 0; i < n_pending; i++) {
          const int p = pend[i];
          T.sortbuf[i].key = ((uint32_t)cnt[p] << 12) | (uint32_t)(uint16_t)box[p].ulx;
          T.sortbuf[i].id = (uint32_t)p;
        }
        int stack[kSortStack];
        std_sort_emulate(T.sortbuf, n_pending, stack);
        int size = S, t = 0;
        for (int j = n_pending - 1; j >= 0; j--) {
          const int p = (int)T.sortbuf[j].id;
          const int ne = (child[4 * p] > 0) + (child[4 * p + 1] > 0) + (child[4 * p + 2] > 0) + (child[4 * p + 3] > 0);
          T.committed[p] = 1;
          T.rank2pos[t++] = (uint16_t)p;
          size += ne - 1;
          if (size >= N) break;
        }
        T.vars[0] = t;
      }
      ORBX_WSYNC();
      n_parents = T.vars[0];
      ORBX_WSYNC();
    }
    // ---- children positions: creation rank r -> first child index; list = reverse(children) ++ survivors ----
    ORBX_LANES(r, n_parents) {
      const int p = T.rank2pos[r];
      T.scan[r] = (child[4 * p] > 0) + (child[4 * p + 1] > 0) + (child[4 * p + 2] > 0) + (child[4 * p + 3] > 0);
    }
    ORBX_WSYNC();
    K = excl_scan(T.scan, n_parents);
    // children boxes / counts, written at their final position; remember who may be split next
    int* child_nxt = T.child[nxt];
    ORBX_LANES(r, n_parents) {
      const int p = T.rank2pos[r];
      int k = T.scan[r];
      const QBox pb = box[p];
      for (int q = 0; q < 4; q++) {
        const int n = child[4 * p + q];
        if (n > 0) {
          const int pos = K - 1 - k;
          T.box[nxt][pos] = child_box(pb, q);
          T.cnt[nxt][pos] = n;
          T.childpos[4 * p + q] = (uint16_t)pos;
          k++;
        }
      }
    }
    ORBX_WSYNC();
    // survivors keep their relative order behind the children
    ORBX_LANES(p, S) T.scan[p] = T.committed[p] ? 0 : 1;
    ORBX_WSYNC();
    const int n_surv = excl_scan(T.scan, S);
    ORBX_LANES(p, S) {
      if (!T.committed[p]) {
        const int pos = K + T.scan[p];
        T.box[nxt][pos] = box[p];
        T.cnt[nxt][pos] = cnt[p];
        T.newpos[p] = (uint16_t)pos;
      }
    }
    ORBX_WSYNC();
    S = K + n_surv;
    // pending list of the next phase-2 round = children with > 1 keys, in creation order (:636-665, :700-720)
    ORBX_LANES(r, n_parents) {
      const int p = T.rank2pos[r];
      T.scan[r] = (child[4 * p] > 1) + (child[4 * p + 1] > 1) + (child[4 * p + 2] > 1) + (child[4 * p + 3] > 1);
    }
    ORBX_WSYNC();
    const int n_expand = excl_scan(T.scan, n_parents);
    ORBX_LANES(r, n_parents) {
      const int p = T.rank2pos[r];
      int k = T.scan[r];
      for (int q = 0; q < 4; q++)
        if (child[4 * p + q] > 1) T.pending[nxt][k++] = T.childpos[4 * p + q];
    }
    ORBX_WSYNC();
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
      ORBX_LANES(p, S) {
        child_nxt[4 * p + 0] = 0;
        child_nxt[4 * p + 1] = 0;
        child_nxt[4 * p + 2] = 0;
        child_nxt[4 * p + 3] = 0;
      }
    } else {
      // final round: child_nxt[p] doubles as "best candidate" accumulator (response << 24 | ~index)
      ORBX_LANES(p, S) child_nxt[p] = 0;
    }
    ORBX_WSYNC();
    // ---- one sweep over the candidates: move to the new node, then histogram / argmax. The loads of 4 lane steps
    //      are issued before any of them is used (memory-level parallelism: a lone warp cannot hide L2 latency) ----
    for (int c0 = ORBX_LANE(); c0 < C; c0 += 4 * ORBX_NLANES) {
      uint32_t lv[4], cv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = c0 + u * ORBX_NLANES;
        if (c < C) {
          lv[u] = T.lab[c];
          cv[u] = T.cand[c];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = c0 + u * ORBX_NLANES;
        if (c >= C) continue;
        const uint32_t l = lv[u];
        const int p = (int)(l & 0xffff);
        const int np = T.committed[p] ? T.childpos[4 * p + (int)(l >> 16)] : T.newpos[p];
        uint32_t nl = (uint32_t)np;
        const uint32_t cw = cv[u];
        if (finish) {
          const uint32_t v = ((uint32_t)cand_s(cw) << 24) | (0xffffffu - (uint32_t)c);
          ORBX_ATOMIC_MAX((unsigned int*)&child_nxt[np], v);
        } else if (split_nxt[np]) {
          const int q = quadrant_of(cw, T.box[nxt][np]);
          nl |= (uint32_t)q << 16;
          ORBX_ATOMIC_ADD(&child_nxt[4 * np + q], 1);
        }
        T.lab[c] = nl;
      }
    }
    ORBX_WSYNC();
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
