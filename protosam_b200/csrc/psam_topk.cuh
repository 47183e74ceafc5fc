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

}  // namespace psam
