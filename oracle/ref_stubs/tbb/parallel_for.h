// Serial stand-in for the two TBB entry points the reference's extractor uses: a range is handed to the body once, in
// order — the serial execution that the oracle defines as the reference result (the TBB twin of ComputeKeyPointsOctTree
// appends to one vector from several tasks, SURVEY.md finding 2; any real schedule is a permutation of this one).
#ifndef ORBREF_STUB_TBB_H_
#define ORBREF_STUB_TBB_H_
namespace tbb {
template <typename T>
struct blocked_range {
  T b, e;
  blocked_range(T b_, T e_) : b(b_), e(e_) {}
  T begin() const { return b; }
  T end() const { return e; }
};
template <typename R, typename F>
void parallel_for(const R& r, const F& f) { f(r); }
template <typename R, typename V, typename F, typename J>
V parallel_reduce(const R& r, const V& init, const F& f, const J&) { return (V)f(r, init); }
}  // namespace tbb
#endif
