// torch.topk(v, 1) on the device, shared by kernel 3b and the optional prompt variants.
#pragma once

namespace psam {

// torch.topk(v, 1) for n < 64: libstdc++ nth_element replayed (oracle: psamo_topk1_pos).
struct TK { float v; int i; };
__device__ __forceinline__ void tk_swap(TK& a, TK& b) { TK t = a; a = b; b = t; }

static __device__ int topk1_small(TK* q, int n)
{
    int first = 0, last = n;
    int depth = 0;
    for (int m = n; m > 1; m >>= 1) ++depth;
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) {
            for (int i = first + 1; i < last; ++i)
                if (q[i].v > q[first].v) tk_swap(q[i], q[first]);
            return q[0].i;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        {   // __move_median_to_first(first, first+1, mid, last-1)
            TK &r = q[first], &a = q[first + 1], &b = q[mid], &c = q[last - 1];
            if (a.v > b.v) {
                if (b.v > c.v) tk_swap(r, b);
                else if (a.v > c.v) tk_swap(r, c);
                else tk_swap(r, a);
            } else if (a.v > c.v) tk_swap(r, a);
            else if (b.v > c.v) tk_swap(r, c);
            else tk_swap(r, b);
        }
        int f = first + 1, l = last;   // __unguarded_partition(first+1, last, pivot=first)
        for (;;) {
            while (q[f].v > q[first].v) ++f;
            --l;
            while (q[first].v > q[l].v) --l;
            if (!(f < l)) break;
            tk_swap(q[f], q[l]);
            ++f;
        }
        if (f <= 0) first = f; else last = f;   // nth == position 0
    }
    // __insertion_sort(first, last)
    for (int i = first + 1; i < last; ++i) {
        TK val = q[i];
        if (val.v > q[first].v) {
            for (int k = i; k > first; --k) q[k] = q[k - 1];
            q[first] = val;
        } else {
            int cur = i, next = i - 1;
            while (val.v > q[next].v) { q[cur] = q[next]; cur = next; --next; }
            q[cur] = val;
        }
    }
    return q[0].i;
}


// ---- torch.topk(v, k) for any k (get_most_conf_points with k > 1, models/ProtoSAM.py:266-289) ----
// ATen's CPU loop: k * 64 <= n -> std::partial_sort(begin, begin + k, end, greater); else std::nth_element(begin,
// begin + k - 1, end, greater) + std::sort(begin, begin + k - 1, greater).  Only values are compared, so where equal
// probabilities land is a function of libstdc++'s algorithms: they are replayed here move for move by one thread
// (oracle: psamo_topk_pos, pinned against torch.topk itself in tests/test_oracle_vs_libs.py).
__device__ __forceinline__ bool tk_gt(const TK& a, const TK& b) { return a.v > b.v; }

static __device__ void tk_push_heap(TK* first, int hole, int top, TK value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && tk_gt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static __device__ void tk_adjust_heap(TK* first, int hole, int len, TK value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (tk_gt(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    tk_push_heap(first, hole, top, value);
}

static __device__ void tk_make_heap(TK* first, int len)
{
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; --parent) {
        const TK value = first[parent];
        tk_adjust_heap(first, parent, len, value);
        if (parent == 0) return;
    }
}

// __pop_heap(first, first + len, result): *result takes the heap's top, the old *result is sifted in
static __device__ void tk_pop_heap(TK* first, int len, TK* result)
{
    const TK value = *result;
    *result = first[0];
    tk_adjust_heap(first, 0, len, value);
}

static __device__ void tk_heap_select(TK* first, int middle, int last)
{
    tk_make_heap(first, middle);
    for (int i = middle; i < last; ++i)
        if (tk_gt(first[i], first[0])) tk_pop_heap(first, middle, first + i);
}

static __device__ void tk_sort_heap(TK* first, int len)
{
    while (len > 1) {
        --len;
        tk_pop_heap(first, len, first + len);
    }
}

static __device__ void tk_insertion_sort(TK* q, int first, int last)
{
    for (int i = first + 1; i < last; ++i) {
        const TK val = q[i];
        if (tk_gt(val, q[first])) {
            for (int k = i; k > first; --k) q[k] = q[k - 1];
            q[first] = val;
        } else {
            int cur = i, next = i - 1;
            while (tk_gt(val, q[next])) { q[cur] = q[next]; cur = next; --next; }
            q[cur] = val;
        }
    }
}

// __unguarded_partition_pivot(first, last)
static __device__ int tk_partition_pivot(TK* q, int first, int last)
{
    const int mid = first + (last - first) / 2;
    TK &r = q[first], &a = q[first + 1], &b = q[mid], &c = q[last - 1];
    if (tk_gt(a, b)) {
        if (tk_gt(b, c)) tk_swap(r, b);
        else if (tk_gt(a, c)) tk_swap(r, c);
        else tk_swap(r, a);
    } else if (tk_gt(a, c)) tk_swap(r, a);
    else if (tk_gt(b, c)) tk_swap(r, c);
    else tk_swap(r, b);
    int f = first + 1, l = last;
    for (;;) {
        while (tk_gt(q[f], q[first])) ++f;
        --l;
        while (tk_gt(q[first], q[l])) --l;
        if (!(f < l)) return f;
        tk_swap(q[f], q[l]);
        ++f;
    }
}

__device__ __forceinline__ int tk_lg(int n) { int d = 0; for (; n > 1; n >>= 1) ++d; return d; }

static __device__ void tk_introselect(TK* q, int first, int nth, int last)
{
    int depth = 2 * tk_lg(last - first);
    while (last - first > 3) {
        if (depth == 0) {
            tk_heap_select(q + first, nth + 1 - first, last - first);
            tk_swap(q[first], q[nth]);
            return;
        }
        --depth;
        const int cut = tk_partition_pivot(q, first, last);
        if (cut <= nth) first = cut; else last = cut;
    }
    tk_insertion_sort(q, first, last);
}

// std::sort(q, q + n): the recursion of __introsort_loop becomes an explicit stack (its sub-ranges are disjoint, so the
// order in which they are finished does not matter); n <= TOPK_MAX_K here
constexpr int TOPK_MAX_K = 64;
static __device__ void tk_sort(TK* q, int n)
{
    if (n <= 0) return;
    int sf[16], sl[16], sd[16], sp = 0;
    sf[0] = 0; sl[0] = n; sd[0] = 2 * tk_lg(n); sp = 1;
    while (sp > 0) {
        --sp;
        int first = sf[sp], last = sl[sp], depth = sd[sp];
        while (last - first > 16) {
            if (depth == 0) {
                tk_heap_select(q + first, last - first, last - first);
                tk_sort_heap(q + first, last - first);
                break;
            }
            --depth;
            const int cut = tk_partition_pivot(q, first, last);
            sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;
            last = cut;
        }
    }
    if (n > 16) {
        tk_insertion_sort(q, 0, 16);
        for (int i = 16; i < n; ++i) {                 // __unguarded_linear_insert
            const TK val = q[i];
            int cur = i, next = i - 1;
            while (tk_gt(val, q[next])) { q[cur] = q[next]; cur = next; --next; }
            q[cur] = val;
        }
    } else {
        tk_insertion_sort(q, 0, n);
    }
}

}  // namespace psam
