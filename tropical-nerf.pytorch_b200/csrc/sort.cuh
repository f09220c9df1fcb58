// sort.cuh -- in-place ascending sort of one short array by ONE thread (registers, local, shared
// or global memory; optional element stride).  Rows are usually 3-8 long (insertion sort); the
// reference's chunk-overlap duplicates produce a few rows of hundreds of coincident vertices,
// and a quadratic sort in one lane then becomes the critical path of the whole kernel: beyond
// kInsertionMax elements heap sort takes over (n log n, no extra storage).
#pragma once

namespace tnb {

constexpr int kInsertionMax = 24;

template <class T>
__device__ __forceinline__ void thread_sort(T *a, int n, int stride = 1)
{
    if (n <= kInsertionMax) {
        for (int i = 1; i < n; ++i) {
            const T key = a[(size_t)i * stride];
            int j = i - 1;
            while (j >= 0 && a[(size_t)j * stride] > key) { a[(size_t)(j + 1) * stride] = a[(size_t)j * stride]; --j; }
            a[(size_t)(j + 1) * stride] = key;
        }
        return;
    }
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        const T v = a[(size_t)root * stride];
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            T cv = a[(size_t)child * stride];
            if (child + 1 < n) {
                const T c2 = a[(size_t)(child + 1) * stride];
                if (cv < c2) { cv = c2; ++child; }
            }
            if (!(v < cv)) break;
            a[(size_t)root * stride] = cv;
            root = child;
        }
        a[(size_t)root * stride] = v;
    }
    for (int end = n - 1; end > 0; --end) {
        const T v = a[(size_t)end * stride];
        a[(size_t)end * stride] = a[0];
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            T cv = a[(size_t)child * stride];
            if (child + 1 < end) {
                const T c2 = a[(size_t)(child + 1) * stride];
                if (cv < c2) { cv = c2; ++child; }
            }
            if (!(v < cv)) break;
            a[(size_t)root * stride] = cv;
            root = child;
        }
        a[(size_t)root * stride] = v;
    }
}

}  // namespace tnb
