// orbx_sort.h — libstdc++'s std::sort reproduced EXACTLY on the device (where it leaves equivalent elements included), single-thread
// and block-parallel.  Used wherever the reference sorts with a comparator that has ties and then depends on the order:
// DistributeOctTree's sort of (count, UL.x) (ORBextractor.cpp:700, k_quadtree.cu) and the frontend's sort of unmatched features by
// response (frontend.cpp:1201-1202, k_cull.cu).
#pragma once
#include <cuda_runtime.h>

// ---- libstdc++ std::sort restated (introsort, threshold 16) on (cnt, ulx, payload) triples ----
// one element = one 64-bit word: count:32 | UL.x:16 | node index:16, so a comparison of (count, UL.x) is one shifted compare
// and every move of the single-thread sort is one LDS/STS.64
typedef unsigned long long SrtE;
__device__ __forceinline__ SrtE srt_make(unsigned cnt, int ulx, int pay) { return ((SrtE)cnt << 32) | ((SrtE)(ulx & 0xFFFF) << 16) | (SrtE)(pay & 0xFFFF); }
__device__ __forceinline__ int srt_pay(SrtE e) { return (int)(e & 0xFFFFu); }
__device__ __forceinline__ bool srt_less(const SrtE &a, const SrtE &b) { return (a >> 16) < (b >> 16); }
__device__ __forceinline__ void srt_swap(SrtE *a, SrtE *b) { SrtE t = *a; *a = *b; *b = t; }
static __device__ void srt_adjust_heap(SrtE *first, int hole, int len, SrtE value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (srt_less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && srt_less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static __device__ void srt_heapsort(SrtE *first, int len)
{
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            SrtE v = first[parent];
            srt_adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        SrtE v = first[last];
        first[last] = first[0];
        srt_adjust_heap(first, 0, last, v);
    }
}
static __device__ void srt_unguarded_linear_insert(SrtE *a, int last)
{
    SrtE val = a[last];
    int next = last - 1;
    while (srt_less(val, a[next])) { a[last] = a[next]; last = next; --next; }
    a[last] = val;
}
static __device__ void srt_insertion_sort(SrtE *a, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (srt_less(a[i], a[first])) {
            SrtE val = a[i];
            for (int j = i; j > first; --j) a[j] = a[j - 1];
            a[first] = val;
        } else srt_unguarded_linear_insert(a, i);
    }
}
static __device__ void srt_sort(SrtE *a, int n)
{
    if (n <= 0) return;
    int lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    // explicit stack replaces the recursion of __introsort_loop(cut, last, depth)
    int stk_first[64], stk_last[64], stk_depth[64], sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = lg * 2; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { srt_heapsort(a + first, last - first); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first+1, mid, last-1)
                SrtE *r = a + first, *x = a + first + 1, *y = a + mid, *z = a + last - 1;
                if (srt_less(*x, *y)) {
                    if (srt_less(*y, *z)) srt_swap(r, y);
                    else if (srt_less(*x, *z)) srt_swap(r, z);
                    else srt_swap(r, x);
                } else if (srt_less(*x, *z)) srt_swap(r, x);
                else if (srt_less(*y, *z)) srt_swap(r, z);
                else srt_swap(r, y);
            }
            int lo = first + 1, hi = last;
            for (;;) {      // __unguarded_partition(first+1, last, pivot = first)
                while (srt_less(a[lo], a[first])) ++lo;
                --hi;
                while (srt_less(a[first], a[hi])) --hi;
                if (!(lo < hi)) break;
                srt_swap(a + lo, a + hi);
                ++lo;
            }
            // recurse on [lo, last) first (the reference recursion), then continue with [first, lo):
            // the right part is fully processed before the left continues, but the two ranges are
            // disjoint, so deferring the right part on a stack gives the same final array.
            stk_first[sp] = lo; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = lo;
        }
    }
    if (n > 16) {
        srt_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) srt_unguarded_linear_insert(a, i);
    } else srt_insertion_sort(a, 0, n);
}

// ---- the same std::sort, block-parallel ----
// libstdc++'s std::sort = __introsort_loop (median-of-3 pivot to the front, Hoare-style __unguarded_partition, recursion until a range
// has <= 16 elements or the depth limit triggers heapsort) + __final_insertion_sort.  Two facts make it parallel WITHOUT changing where
// it leaves equivalent elements:
//   * a partition is determined by the ORIGINAL values of its range: the left scan stops at the positions whose element is not below the
//     pivot (ascending), the right scan at those not above it (descending); the k-th stops of both sides are swapped while they have not
//     crossed, and the cut is the first left stop after the last swap (or the last right stop if none lies before it).  So one warp does
//     a whole partition with two ballot scans, a monotone count and independent swaps; disjoint ranges of one recursion depth go to
//     different warps;
//   * insertion sort is stable, so the final pass equals a stable sort of whatever the partitions left: a rank computation.
// Checked against the single-thread restatement above (and through it against the real std::sort) on tie-heavy inputs.
__device__ __forceinline__ void srt_median_to_first(SrtE *a, int first, int mid, int last)
{
    SrtE *r = a + first, *x = a + first + 1, *y = a + mid, *z = a + last - 1;
    if (srt_less(*x, *y)) {
        if (srt_less(*y, *z)) srt_swap(r, y);
        else if (srt_less(*x, *z)) srt_swap(r, z);
        else srt_swap(r, x);
    } else if (srt_less(*x, *z)) srt_swap(r, x);
    else if (srt_less(*y, *z)) srt_swap(r, z);
    else srt_swap(r, y);
}

// __unguarded_partition(lo, hi, pivot) by one warp; Ls / Ra are scratch slices of at least hi - lo entries.  Returns the cut.
static __device__ int srt_warp_partition(SrtE *a, int lo, int hi, SrtE pivot, unsigned short *Ls, unsigned short *Ra)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int nL = 0, nR = 0;
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const bool in = i < hi;
        const SrtE v = in ? a[i] : 0ull;
        const bool fl = in && !srt_less(v, pivot), fr = in && !srt_less(pivot, v);
        const unsigned bl = __ballot_sync(0xffffffffu, fl), br = __ballot_sync(0xffffffffu, fr);
        if (fl) Ls[nL + __popc(bl & lt)] = (unsigned short)i;
        if (fr) Ra[nR + __popc(br & lt)] = (unsigned short)i;          // ascending; the k-th right stop is Ra[nR - 1 - k]
        nL += __popc(bl); nR += __popc(br);
    }
    __syncwarp();
    const int nmin = min(nL, nR);
    int K = 0;                                                           // number of swaps: stops cross monotonically
    for (int base = 0; base < nmin; base += 32) {
        const int k = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, k < nmin && Ls[k] < Ra[nR - 1 - k]);
        K += __popc(b);
        if (b != 0xffffffffu) break;
    }
    int cut;
    if (K < nL && (K == 0 || Ls[K] < Ra[nR - K])) cut = Ls[K]; else cut = Ra[nR - K];
    for (int k = lane; k < K; k += 32) srt_swap(a + Ls[k], a + Ra[nR - 1 - k]);
    __syncwarp();
    return cut;
}

// all NT threads; a[0..n) in shared memory.  tmp: n elements; Ls, Ra: n entries each; lists: 4*cap words; cnt: 2 ints.
template <int NT> __device__ void block_sort_exact(SrtE *a, int n, SrtE *tmp, unsigned short *Ls, unsigned short *Ra, unsigned *lists, int cap, int *cnt)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (n > 16) {
        if (threadIdx.x == 0) {
            int lg = 0;
            while ((n >> (lg + 1)) > 0) lg++;
            lists[0] = (unsigned)n << 16; lists[1] = (unsigned)(2 * lg); cnt[0] = 1; cnt[1] = 0;
        }
        __syncthreads();
        unsigned *cur = lists, *nxt = lists + 2 * cap;
        while (true) {
            const int nr = cnt[0];
            if (nr == 0) break;
            for (int r = wid; r < nr; r += NT / 32) {
                const int first = (int)(cur[2 * r] & 0xFFFFu), last = (int)(cur[2 * r] >> 16);
                int depth = (int)cur[2 * r + 1];
                if (depth == 0) { if (lane == 0) srt_heapsort(a + first, last - first); __syncwarp(); continue; }
                depth--;
                if (lane == 0) srt_median_to_first(a, first, first + (last - first) / 2, last);
                __syncwarp();
                const int cut = srt_warp_partition(a, first + 1, last, a[first], Ls + first, Ra + first);
                if (lane == 0) {
                    if (last - cut > 16) { const int sl = atomicAdd(&cnt[1], 1); nxt[2 * sl] = (unsigned)cut | ((unsigned)last << 16); nxt[2 * sl + 1] = (unsigned)depth; }
                    if (cut - first > 16) { const int sl = atomicAdd(&cnt[1], 1); nxt[2 * sl] = (unsigned)first | ((unsigned)cut << 16); nxt[2 * sl + 1] = (unsigned)depth; }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) { cnt[0] = cnt[1]; cnt[1] = 0; }
            unsigned *t = cur; cur = nxt; nxt = t;
            __syncthreads();
        }
    }
    // __final_insertion_sort == stable sort of the current arrangement
    for (int i = threadIdx.x; i < n; i += NT) {
        const SrtE v = a[i];
        const unsigned long long key = v >> 16;
        int rank = 0;
        for (int j = 0; j < n; j++) { const unsigned long long kj = a[j] >> 16; rank += (kj < key) || (kj == key && j < i); }
        tmp[rank] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) a[i] = tmp[i];
    __syncthreads();
}

