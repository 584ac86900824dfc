// kp_gosort.cuh -- Go's sort.Slice (pdqsort_func of package sort, go1.26) on the claim-order arrays, device side.
#pragma once
#include <cuda_runtime.h>

// ---- Go's sort.Slice (pdqsort_func, package sort of go1.26) on the claim-order arrays -------------------------
// Emulated exactly because the permutation it leaves among claims with EQUAL pod counts decides first-fit
// (scheduler.go:504).  key = len(Pods) by position, val = claim id by position.
struct DevSorter {
  int* key;
  int* val;
  __device__ bool less(int i, int j) const { return key[i] < key[j]; }
  __device__ void swap(int i, int j) {
    int t = key[i];
    key[i] = key[j];
    key[j] = t;
    t = val[i];
    val[i] = val[j];
    val[j] = t;
  }
  __device__ void insertion_sort(int a, int b) {
    for (int i = a + 1; i < b; i++)
      for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
  }
  __device__ void sift_down(int lo, int hi, int first) {
    int root = lo;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  __device__ void heap_sort(int a, int b) {
    int first = a, lo = 0, hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) sift_down(i, hi, first);
    for (int i = hi - 1; i >= 0; i--) {
      swap(first, first + i);
      sift_down(lo, i, first);
    }
  }
  __device__ int partition(int a, int b, int pivot, bool* already) {
    swap(a, pivot);
    int i = a + 1, j = b - 1;
    while (i <= j && less(i, a)) i++;
    while (i <= j && !less(j, a)) j--;
    if (i > j) {
      swap(j, a);
      *already = true;
      return j;
    }
    swap(i, j);
    i++;
    j--;
    for (;;) {
      while (i <= j && less(i, a)) i++;
      while (i <= j && !less(j, a)) j--;
      if (i > j) break;
      swap(i, j);
      i++;
      j--;
    }
    swap(j, a);
    *already = false;
    return j;
  }
  __device__ int partition_equal(int a, int b, int pivot) {
    swap(a, pivot);
    int i = a + 1, j = b - 1;
    for (;;) {
      while (i <= j && !less(a, i)) i++;
      while (i <= j && less(a, j)) j--;
      if (i > j) break;
      swap(i, j);
      i++;
      j--;
    }
    return i;
  }
  __device__ bool partial_insertion_sort(int a, int b) {
    int i = a + 1;
    for (int j = 0; j < 5; j++) {
      while (i < b && !less(i, i - 1)) i++;
      if (i == b) return true;
      if (b - a < 50) return false;
      swap(i, i - 1);
      if (i - a >= 2)
        for (int k = i - 1; k >= 1; k--) {
          if (!less(k, k - 1)) break;
          swap(k, k - 1);
        }
      if (b - i >= 2)
        for (int k = i + 1; k < b; k++) {
          if (!less(k, k - 1)) break;
          swap(k, k - 1);
        }
    }
    return false;
  }
  __device__ static int bits_len(unsigned long long x) { return x ? 64 - __clzll((long long)x) : 0; }
  __device__ void break_patterns(int a, int b) {
    int length = b - a;
    if (length >= 8) {
      unsigned long long random = (unsigned long long)length;
      unsigned long long modulus = 1ull << bits_len((unsigned long long)length);
      int idx = a + (length / 4) * 2 - 1;
      for (int i = 0; i < 3; i++) {
        random ^= random << 13;
        random ^= random >> 7;
        random ^= random << 17;
        int other = (int)(random & (modulus - 1));
        if (other >= length) other -= length;
        swap(idx - 1 + i, a + other);
      }
    }
  }
  __device__ void order2(int* a, int* b, int* swaps) const {
    if (less(*b, *a)) {
      (*swaps)++;
      int t = *a;
      *a = *b;
      *b = t;
    }
  }
  __device__ int median(int a, int b, int c, int* swaps) const {
    order2(&a, &b, swaps);
    order2(&b, &c, swaps);
    order2(&a, &b, swaps);
    return b;
  }
  // returns pivot; hint: 0 unknown, 1 increasing, 2 decreasing
  __device__ int choose_pivot(int a, int b, int* hint) const {
    int l = b - a, swaps = 0;
    int i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
    if (l >= 8) {
      if (l >= 50) {
        i = median(i - 1, i, i + 1, &swaps);
        j = median(j - 1, j, j + 1, &swaps);
        k = median(k - 1, k, k + 1, &swaps);
      }
      j = median(i, j, k, &swaps);
    }
    *hint = swaps == 0 ? 1 : (swaps == 12 ? 2 : 0);
    return j;
  }
  __device__ void reverse_range(int a, int b) {
    int i = a, j = b - 1;
    while (i < j) {
      swap(i, j);
      i++;
      j--;
    }
  }
  __device__ void pdqsort(int a, int b, int limit) {
    bool wasBalanced = true, wasPartitioned = true;
    for (;;) {
      int length = b - a;
      if (length <= 12) {
        insertion_sort(a, b);
        return;
      }
      if (limit == 0) {
        heap_sort(a, b);
        return;
      }
      if (!wasBalanced) {
        break_patterns(a, b);
        limit--;
      }
      int hint;
      int pivot = choose_pivot(a, b, &hint);
      if (hint == 2) {
        reverse_range(a, b);
        pivot = (b - 1) - (pivot - a);
        hint = 1;
      }
      if (wasBalanced && wasPartitioned && hint == 1) {
        if (partial_insertion_sort(a, b)) return;
      }
      if (a > 0 && !less(a - 1, pivot)) {
        a = partition_equal(a, b, pivot);
        continue;
      }
      bool already;
      int mid = partition(a, b, pivot, &already);
      wasPartitioned = already;
      int leftLen = mid - a, rightLen = b - mid;
      int balanceThreshold = length / 8;
      if (leftLen < rightLen) {
        wasBalanced = leftLen >= balanceThreshold;
        pdqsort(a, mid, limit);
        a = mid + 1;
      } else {
        wasBalanced = rightLen >= balanceThreshold;
        pdqsort(mid + 1, b, limit);
        b = mid;
      }
    }
  }
};


