// kp_gosort_host.hpp -- Go's sort.Slice (pdqsort_func, package sort of go1.19 .. go1.26) on the host, for the
// order-sensitive sorts that stay on the CPU side of the boundary: sortCandidates by DisruptionCost
// (pkg/controllers/disruption/consolidation.go:126-131, singlenodeconsolidation.go:143-146).  sort.Slice is not stable;
// which of two equally expensive candidates comes first decides the prefix the multi-node search evaluates, so the tie
// permutation is reproduced, not approximated.  (The device twin -- warp-cooperative, for the per-pod NodeClaim sort -- is
// kp_gosort.cuh.)  Scalar and self-contained: Less / Swap act on a permutation of the keys.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

template <class Key>
struct HostGoSort {
  const Key* key;
  int32_t* perm;  // perm[i] = index of the element currently at position i

  bool less(int i, int j) const { return key[perm[i]] < key[perm[j]]; }
  void swap(int i, int j) { std::swap(perm[i], perm[j]); }
  static int bits_len(unsigned long long x) {
    int n = 0;
    for (; x; x >>= 1) n++;
    return n;
  }

  void insertion_sort(int a, int b) {
    for (int i = a + 1; i < b; i++)
      for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
  }
  void sift_down(int lo, int hi, int first) {
    for (int root = lo;;) {
      int child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  void heap_sort(int a, int b) {
    const int first = a, hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) sift_down(i, hi, first);
    for (int i = hi - 1; i >= 0; i--) {
      swap(first, first + i);
      sift_down(0, i, first);
    }
  }
  int order2(int a, int b, int* swaps, int* hi) {  // returns the smaller position, *hi the larger
    if (less(b, a)) {
      (*swaps)++;
      *hi = a;
      return b;
    }
    *hi = b;
    return a;
  }
  int median(int a, int b, int c, int* swaps) {
    int t;
    a = order2(a, b, swaps, &t);
    b = t;
    b = order2(b, c, swaps, &t);
    c = t;
    a = order2(a, b, swaps, &t);
    b = t;
    (void)a;
    (void)c;
    return b;
  }
  // 0 unknown, 1 increasing, 2 decreasing
  int choose_pivot(int a, int b, int* hint) {
    const int l = b - a;
    int swaps = 0, i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
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
  void reverse_range(int a, int b) {
    for (int i = a, j = b - 1; i < j; i++, j--) swap(i, j);
  }
  bool partial_insertion_sort(int a, int b) {
    int i = a + 1;
    for (int step = 0; step < 5; step++) {
      while (i < b && !less(i, i - 1)) i++;
      if (i == b) return true;
      if (b - a < 50) return false;
      swap(i, i - 1);
      if (i - a >= 2)
        for (int j = i - 1; j >= 1; j--) {
          if (!less(j, j - 1)) break;
          swap(j, j - 1);
        }
      if (b - i >= 2)
        for (int j = i + 1; j < b; j++) {
          if (!less(j, j - 1)) break;
          swap(j, j - 1);
        }
    }
    return false;
  }
  void break_patterns(int a, int b) {
    const int length = b - a;
    if (length < 8) return;
    unsigned long long r = (unsigned long long)length;
    const unsigned long long modulus = 1ull << bits_len((unsigned long long)length);
    const int idx = a + (length / 4) * 2 - 1;
    for (int i = 0; i < 3; i++) {
      r ^= r << 13;
      r ^= r >> 7;
      r ^= r << 17;
      int other = (int)(r & (modulus - 1));
      if (other >= length) other -= length;
      swap(idx - 1 + i, a + other);
    }
  }
  int partition_equal(int a, int b, int pivot) {
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
  int partition(int a, int b, int pivot, bool* already) {
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
  void pdqsort(int a, int b, int limit) {
    bool was_balanced = true, was_partitioned = true;
    for (;;) {
      const int length = b - a;
      if (length <= 12) {
        insertion_sort(a, b);
        return;
      }
      if (limit == 0) {
        heap_sort(a, b);
        return;
      }
      if (!was_balanced) {
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
      if (was_balanced && was_partitioned && hint == 1 && partial_insertion_sort(a, b)) return;
      if (a > 0 && !less(a - 1, pivot)) {
        a = partition_equal(a, b, pivot);
        continue;
      }
      bool already;
      const int mid = partition(a, b, pivot, &already);
      was_partitioned = already;
      const int left = mid - a, right = b - mid, threshold = length / 8;
      if (left < right) {
        was_balanced = left >= threshold;
        pdqsort(a, mid, limit);
        a = mid + 1;
      } else {
        was_balanced = right >= threshold;
        pdqsort(mid + 1, b, limit);
        b = mid;
      }
    }
  }
};

// perm_out[i] = index of the element sort.Slice(keys, less = <) leaves at position i
template <class Key>
inline void host_go_sort(const Key* keys, int n, int32_t* perm_out) {
  for (int i = 0; i < n; i++) perm_out[i] = i;
  HostGoSort<Key> s{keys, perm_out};
  s.pdqsort(0, n, HostGoSort<Key>::bits_len((unsigned long long)n));
}
