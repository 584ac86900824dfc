// kp_solve.cuh -- k_solve: the persistent Scheduler.Solve kernel (one CTA == one Scheduler).
#pragma once
#include "kp_kernels.cuh"

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


enum { PERT_NONE = 0, PERT_INC = 1, PERT_APPEND = 2 };

// Everything the per-pod critical path touches lives in shared memory: the claim order (s.newNodeClaims as ids +
// pod counts), per-claim template ids, the "can never fit again" bits, the allocatable thresholds and the staged pod.
struct SolveShared {
  int head, tail, cap;
  int done, found;
  int n_claims;
  int pert_kind, pert_pos;
  int rot_from, rot_to, rot_mode;
  int ctx_idx[2];  // queue index whose class row sits in ctx[i]
  int alive_tmpl;
  unsigned warp_mask[SOLVE_WARPS];
  int ok[SOLVE_WARPS];
  PodCtx ctx[2];
  Slot scratch[SOLVE_WARPS][KP_MAXK];
  KpDev ds;  // the table pointers, patched to the shared-memory copies of the small read-only tables
};

#define KP_ALIGN16(x) (((x) + 15) & ~(size_t)15)
// bytes of the read-only tables k_solve stages in shared memory (same formula on host and device)
__host__ __device__ inline size_t kp_tab_bytes(const KpDev& d) {
  size_t K = d.K, R = d.R, ITW = d.ITW, D = d.D > 0 ? d.D : 1, N = d.N > 0 ? d.N : 1;
  size_t b = 0;
  b += KP_ALIGN16(K) + 2 * KP_ALIGN16(8 * K);                    // key_wellknown, key_univ, val_isint
  b += KP_ALIGN16(4 * (R + 1)) + KP_ALIGN16(8 * (size_t)d.n_ge) + KP_ALIGN16(8 * (size_t)d.n_ge * ITW);
  b += KP_ALIGN16(4 * (K + 1)) + KP_ALIGN16(8 * (size_t)d.n_itv * ITW);
  b += 3 * KP_ALIGN16(8 * K * ITW) + KP_ALIGN16(8 * ITW);        // it_nokey, it_dne, it_nonempty, it_valid
  b += KP_ALIGN16(sizeof(Slot) * D * K) + KP_ALIGN16(4 * D) + KP_ALIGN16(8 * D * ITW);
  b += KP_ALIGN16(4 * N);
  return b;
}

// n-th set bit over the per-warp ballot masks of a chunk (-1 if fewer); *total = number of set bits
__device__ __forceinline__ int nth_candidate(const unsigned* masks, int n, int* total) {
  int acc = 0, found = -1;
#pragma unroll
  for (int w = 0; w < SOLVE_WARPS; w++) {
    unsigned m = masks[w];
    int c = __popc(m);
    if (found < 0 && n < acc + c) found = w * 32 + (__fns(m, 0, n - acc + 1));
    acc += c;
  }
  *total = acc;
  return found;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(const __grid_constant__ KpDev d_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SolveShared& sh = *reinterpret_cast<SolveShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- stage the small read-only tables (key universe, thresholds, bit-sliced instance-type rows, offering sets)
  // in shared memory and keep a patched copy of the pointer block next to them
  {
    const int* src = reinterpret_cast<const int*>(&d_in);
    int* dst = reinterpret_cast<int*>(&sh.ds);
    for (int i = tid; i < (int)(sizeof(KpDev) / 4); i += SOLVE_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  unsigned char* tab = smem_raw + KP_ALIGN16(sizeof(SolveShared));
  if (d_in.tab_bytes > 0) {
    size_t off = 0;
#define KP_STAGE(field, type, count)                                                            \
  {                                                                                             \
    size_t n_ = (size_t)(count);                                                                \
    type* dst_ = reinterpret_cast<type*>(tab + off);                                            \
    const type* src_ = d_in.field;                                                              \
    for (size_t i_ = tid; i_ < n_; i_ += SOLVE_THREADS) dst_[i_] = src_[i_];                    \
    if (tid == 0) sh.ds.field = dst_;                                                           \
    off += KP_ALIGN16(sizeof(type) * n_);                                                       \
  }
    const size_t K_ = d_in.K, R_ = d_in.R, W_ = d_in.ITW, D_ = d_in.D > 0 ? d_in.D : 1, N_ = d_in.N > 0 ? d_in.N : 1;
    KP_STAGE(key_wellknown, uint8_t, K_)
    KP_STAGE(key_univ, uint64_t, K_)
    KP_STAGE(val_isint, uint64_t, K_)
    KP_STAGE(ge_off, int32_t, R_ + 1)
    KP_STAGE(ge_vals, int64_t, d_in.n_ge)
    KP_STAGE(ge_bits, uint64_t, (size_t)d_in.n_ge * W_)
    KP_STAGE(itv_off, int32_t, K_ + 1)
    KP_STAGE(itv, uint64_t, (size_t)d_in.n_itv * W_)
    KP_STAGE(it_nokey, uint64_t, K_ * W_)
    KP_STAGE(it_dne, uint64_t, K_ * W_)
    KP_STAGE(it_nonempty, uint64_t, K_ * W_)
    KP_STAGE(it_valid, uint64_t, W_)
    KP_STAGE(off_slots, Slot, D_ * K_)
    KP_STAGE(off_keys, uint32_t, D_)
    KP_STAGE(offset_bits, uint64_t, D_ * W_)
    KP_STAGE(tmpl_taintset, int32_t, N_)
#undef KP_STAGE
  }
  __syncthreads();
  const KpDev& d = sh.ds;
  const int K = d.K, R = d.R, ITW = d.ITW, E = d.E, CS = d.CS;
  // shared mirrors (write-through) of the first CS claim positions / ids
  int* order_s = reinterpret_cast<int*>(tab + d_in.tab_bytes);
  int* cnt_s = order_s + CS;
  int* tmpl_s = cnt_s + CS;
  uint32_t* rdead_s = reinterpret_cast<uint32_t*>(tmpl_s + CS);  // [n_rv * RWS]
  const int RW = (d.Cmax + 31) >> 5, RWS = (CS + 31) >> 5;
  auto ord = [&](int pos) { return pos < CS ? order_s[pos] : d.order[pos]; };
  auto cnt = [&](int pos) { return pos < CS ? cnt_s[pos] : d.cnt_at[pos]; };
  auto set_ord = [&](int pos, int c, int n) {
    d.order[pos] = c;
    d.cnt_at[pos] = n;
    if (pos < CS) {
      order_s[pos] = c;
      cnt_s[pos] = n;
    }
  };
  auto claim_tmpl = [&](int c) { return c < CS ? tmpl_s[c] : d.c_tmpl[c]; };
  auto is_rdead = [&](int rv, int c) {
    uint32_t w = c < CS ? rdead_s[rv * RWS + (c >> 5)] : d.rdead[(size_t)rv * RW + (c >> 5)];
    return (w >> (c & 31)) & 1u;
  };
  if (tid == 0) {
    sh.head = 0;
    sh.tail = (int)d.P;
    sh.cap = (int)d.P + 1;
    sh.n_claims = 0;
    sh.pert_kind = PERT_NONE;
    sh.done = 0;
    sh.ctx_idx[0] = sh.ctx_idx[1] = -1;
    int alive = 0;
    for (int n = 0; n < d.N; n++) {
      bool any = false;
      for (int w = 0; w < ITW; w++) any |= d.tmpl_its[(size_t)n * ITW + w] != 0;
      alive += any;
    }
    sh.alive_tmpl = alive;
  }
  for (int i = tid; i < d.n_rv * RWS; i += SOLVE_THREADS) rdead_s[i] = 0;
  long long ev_existing = 0, ev_inflight = 0, ev_tmpl = 0, commits = 0, slow_sorts = 0;  // thread 0 only
  int n_active_nodes = 0;
  for (int n = tid; n < E; n += SOLVE_THREADS) n_active_nodes += (d.node_flags[n] & KP_NODE_SCHEDULABLE) ? 1 : 0;
  {
    for (int o = 16; o; o >>= 1) n_active_nodes += __shfl_xor_sync(FULL, n_active_nodes, o);
    __syncthreads();
    if (lane == 0) sh.ok[warp] = n_active_nodes;
    __syncthreads();
    n_active_nodes = 0;
    for (int i = 0; i < SOLVE_WARPS; i++) n_active_nodes += sh.ok[i];
  }
  // class prefetch pipeline of the last warp: pf_cls = class of queue index pf_idx (loaded one iteration earlier)
  int a_idx = -1, a_val = -1;

  long long watchdog = 0;
  long long t_sort = 0, t_inflight = 0, t_new = 0, n_rounds = 0, n_ctx_miss = 0, t_pop = 0, t_mark = 0;
  for (;;) {
    __syncthreads();  // (S0) previous pod fully committed; prefetched rows visible
    const int h = sh.head;
    t_mark = clock64();
    if (++watchdog > 4 * (long long)d.P + 1024) {  // cannot happen: every requeue cycle needs progress (queue.go:54-58)
      if (tid == 0) *d.status = KP_ERR_INVALID;
      break;
    }
    // ---- Queue.Pop (queue.go:46-60)
    if (tid == 0) {
      int len = sh.tail - h;
      if (len == 0) {
        sh.done = 1;
      } else if (h >= (int)d.P && d.last_len[d.queue[h % sh.cap]] == len) {
        sh.done = 1;  // a full cycle without progress
      }
      sh.found = 0;
    }
    // stage the pod's class row unless the prefetcher already did
    const bool have_ctx = sh.ctx_idx[h & 1] == h;
    const int pert = sh.pert_kind;  // read between S0 and S1: commits write it before S0, the sort stage after S1
    __syncthreads();  // (S1)
    if (sh.done) break;
    if (!have_ctx) {
      n_ctx_miss++;
      if (warp == 0) {
        ClassRegs cr = load_class_regs(d, d.qcls[h % sh.cap], d.queue[h % sh.cap], lane);
        store_class_regs(d, sh.ctx[h & 1], cr, lane);
        if (lane == 0) sh.ctx_idx[h & 1] = h;
      }
      __syncthreads();
    }
    if (tid == 0) sh.head = h + 1;
    // ---- software-pipelined staging of upcoming pods by the last warp: stage A (previous iteration) fetched
    // (class, pod) of queue index h+1 into registers; stage B issues that pod's class-row loads now and parks them
    // in registers until the end of this iteration, so no warp ever waits on them.
    int b_idx = -1;
    ClassRegs br;
    if (warp == SOLVE_WARPS - 1) {
      if (a_idx == h + 1) {
        int X = __shfl_sync(FULL, a_val, 0);
        int podn = __shfl_sync(FULL, a_val, 1);
        b_idx = a_idx;
        br = load_class_regs(d, X, podn, lane);
      }
      a_idx = -1;
      int nxt = h + 2;
      if (nxt < sh.tail) {
        a_idx = nxt;
        if (lane == 0) a_val = d.qcls[nxt % sh.cap];
        if (lane == 1) a_val = d.queue[nxt % sh.cap];
      }
    }
    const PodCtx& px = sh.ctx[h & 1];
    const int pod = px.pod;
    const int tolset = px.tolset, rv = px.rv, sig = px.sig;
    do {

    // ================= addToExistingNode (scheduler.go:520-555) =================
    for (int base = 0; base < E && !sh.found; base += SOLVE_THREADS) {
      int n = base + tid;
      bool pass = false;
      if (n < E && (d.node_flags[n] & KP_NODE_SCHEDULABLE)) {
        pass = tolerated(d, tolset, d.node_taintset[n]);
        if (pass) {  // resources.Fits(pod requests, remaining) (resources.go:150-163)
          uint32_t pr = d.node_rem_present[n];
          for (int r = 0; r < R; r++) {
            int64_t rem = d.node_rem[(size_t)n * R + r];
            bool present = pr >> r & 1;
            if (present && rem < 0) pass = false;
            if (px.req[r] > (present ? rem : 0)) pass = false;
          }
        }
      }
      unsigned m = __ballot_sync(FULL, pass);
      if (lane == 0) sh.warp_mask[warp] = m;
      __syncthreads();
      int total;
      for (int g0 = 0;; g0 += SOLVE_WARPS) {
        int off = nth_candidate(sh.warp_mask, g0 + warp, &total);
        if (g0 >= total) break;
        Eval ev;
        ev.ok = false;
        int node = -1;
        if (off >= 0) {
          node = base + off;
          Slot b = lane < K ? load_slot(d.node_sflags, d.node_smask, d.node_sgte, d.node_slte, (size_t)node * K + lane,
                                         d.has_bounds)
                            : slot_absent();
          ev = eval_candidate(d, px, false, b, 0, 0, node, sh.scratch[warp], lane);
        }
        if (lane == 0) sh.ok[warp] = ev.ok;
        __syncthreads();
        int winner = -1;
#pragma unroll
        for (int i = SOLVE_WARPS - 1; i >= 0; i--)
          if (sh.ok[i]) winner = i;
        if (winner == warp) {  // ExistingNode.Add (existingnode.go:147-155)
          if (lane < K) {
            size_t i = (size_t)node * K + lane;
            d.node_sflags[i] = (uint8_t)ev.F.f;
            d.node_smask[i] = ev.F.m;
            if (d.has_bounds) {
              d.node_sgte[i] = ev.F.gte;
              d.node_slte[i] = ev.F.lte;
            }
          }
          if (lane < R) d.node_rem[(size_t)node * R + lane] -= px.req[lane];
          if (lane == 0) {
            d.node_rem_present[node] |= (1u << R) - 1;
            d.node_npods[node]++;
            d.pod_target[pod] = node;
            d.pod_error[pod] = KP_PODERR_NONE;
            d.counters[0] += node + 1;
            sh.found = 1;
          }
          topo_record(d, px, ev.F, d.node_taintset[node], node, false, lane);
        }
        __syncthreads();
        if (winner >= 0) break;
      }
      __syncthreads();  // the next chunk rewrites warp_mask
    }
    if (tid == 0) {
      if (sh.found)
        commits++;
      else
        ev_existing += n_active_nodes;
    }
    if (sh.found) break;

    // ================= sort.Slice(newNodeClaims, len(Pods) asc) (scheduler.go:504) =================
    t_pop += clock64() - t_mark;
    t_mark = clock64();
    const int nC = sh.n_claims;
    if (pert != PERT_NONE) {
      if (tid == 0) {
        int p = sh.pert_pos;
        bool inversion = sh.pert_kind == PERT_INC ? (p + 1 < nC && cnt(p + 1) < cnt(p))
                                                  : (nC >= 2 && cnt(nC - 1) < cnt(nC - 2));
        sh.rot_mode = 0;
        if (inversion) {
          bool stable = d.stable_order || nC <= 12;
          bool in_smem = nC <= CS;
          if (!stable && nC >= 50) {
            DevSorter s{in_smem ? cnt_s : d.cnt_at, in_smem ? order_s : d.order};
            int hint;
            s.choose_pivot(0, nC, &hint);
            stable = hint == 1;  // partialInsertionSort repairs a single inversion == stable move
          }
          if (stable) {
            if (sh.pert_kind == PERT_INC) {
              int c = cnt(p);  // elevated count; move right past every smaller element
              int lo = p + 1, hi = nC;
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cnt(mid) < c)
                  lo = mid + 1;
                else
                  hi = mid;
              }
              sh.rot_from = p;
              sh.rot_to = lo - 1;
              sh.rot_mode = 1;
            } else {
              int c = cnt(nC - 1);  // new claim: move left past every larger element
              int lo = 0, hi = nC - 1;
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cnt(mid) <= c)
                  lo = mid + 1;
                else
                  hi = mid;
              }
              sh.rot_from = nC - 1;
              sh.rot_to = lo;
              sh.rot_mode = 2;
            }
          } else {
            // exact pdqsort emulation by one thread (rare: ties scrambled by Go's unstable partition)
            DevSorter s{in_smem ? cnt_s : d.cnt_at, in_smem ? order_s : d.order};
            if (!in_smem)
              for (int i = 0; i < CS; i++) {  // global arrays are the source of truth beyond CS
                d.cnt_at[i] = cnt_s[i];
                d.order[i] = order_s[i];
              }
            s.pdqsort(0, nC, DevSorter::bits_len((unsigned long long)nC));
            if (in_smem) {
              for (int i = 0; i < nC; i++) {
                d.cnt_at[i] = cnt_s[i];
                d.order[i] = order_s[i];
              }
            } else {
              for (int i = 0; i < CS; i++) {
                cnt_s[i] = d.cnt_at[i];
                order_s[i] = d.order[i];
              }
            }
            slow_sorts++;
          }
        }
        sh.pert_kind = PERT_NONE;
      }
      __syncthreads();
      if (sh.rot_mode == 1) {  // rotate [from, to] left by one
        int from = sh.rot_from, to = sh.rot_to;
        int eo = ord(from), ec = cnt(from);
        for (int b0 = from; b0 < to; b0 += SOLVE_THREADS) {
          int i = b0 + tid;
          int vo = 0, vc = 0;
          if (i < to) {
            vo = ord(i + 1);
            vc = cnt(i + 1);
          }
          __syncthreads();
          if (i < to) set_ord(i, vo, vc);
          __syncthreads();
        }
        if (tid == 0) set_ord(to, eo, ec);
        __syncthreads();
      } else if (sh.rot_mode == 2) {  // rotate [to, from] right by one
        int from = sh.rot_from, to = sh.rot_to;
        int eo = ord(from), ec = cnt(from);
        for (int b0 = from; b0 > to; b0 -= SOLVE_THREADS) {
          int i = b0 - tid;
          int vo = 0, vc = 0;
          if (i > to) {
            vo = ord(i - 1);
            vc = cnt(i - 1);
          }
          __syncthreads();
          if (i > to) set_ord(i, vo, vc);
          __syncthreads();
        }
        if (tid == 0) set_ord(to, eo, ec);
        __syncthreads();
      }
    }

    // ================= addToInflightNode (scheduler.go:557-589) =================
    t_sort += clock64() - t_mark;
    t_mark = clock64();
    for (int base = 0; base < nC && !sh.found; base += SOLVE_THREADS) {
      int pos = base + tid;
      bool pass = false;
      if (pos < nC) {
        int c = ord(pos);
        pass = !is_rdead(rv, c) && ((px.tmpl_ok >> claim_tmpl(c)) & 1ull);
        if (pass && sig >= 0) pass = d.fver[(size_t)sig * d.Cmax + c] != d.cver[c] + 1;  // known failure, claim unchanged
      }
      unsigned m = __ballot_sync(FULL, pass);
      if (lane == 0) sh.warp_mask[warp] = m;
      __syncthreads();
      int total;
      for (int g0 = 0;; g0 += SOLVE_WARPS) {
        int off = nth_candidate(sh.warp_mask, g0 + warp, &total);
        if (g0 >= total) break;
        Eval ev;
        ev.ok = false;
        ev.res_dead = false;
        n_rounds++;
        int cpos = -1, cc = -1;
        if (off >= 0) {
          cpos = base + off;
          cc = ord(cpos);
          Slot b = lane < K ? load_slot(d.c_sflags, d.c_smask, d.c_sgte, d.c_slte, (size_t)cc * K + lane, d.has_bounds)
                            : slot_absent();
          int64_t bq = lane < R ? d.c_req[(size_t)cc * R + lane] : 0;
          uint64_t bi = lane < ITW ? d.c_its[(size_t)cc * ITW + lane] : 0ull;
          ev = eval_candidate(d, px, true, b, bq, bi, E + cc, sh.scratch[warp], lane);
          if (ev.res_dead && lane == 0) {
            atomicOr(&d.rdead[(size_t)rv * RW + (cc >> 5)], 1u << (cc & 31));
            if (cc < CS) atomicOr(&rdead_s[rv * RWS + (cc >> 5)], 1u << (cc & 31));
          }
          if (!ev.ok && sig >= 0 && lane == 0) d.fver[(size_t)sig * d.Cmax + cc] = d.cver[cc] + 1;
        }
        if (lane == 0) sh.ok[warp] = ev.ok;
        __syncthreads();
        int winner = -1;
#pragma unroll
        for (int i = SOLVE_WARPS - 1; i >= 0; i--)
          if (sh.ok[i]) winner = i;
        if (winner == warp) {  // NodeClaim.Add (nodeclaim.go:207-219)
          if (lane < K) {
            size_t i = (size_t)cc * K + lane;
            d.c_sflags[i] = (uint8_t)ev.F.f;
            d.c_smask[i] = ev.F.m;
            if (d.has_bounds) {
              d.c_sgte[i] = ev.F.gte;
              d.c_slte[i] = ev.F.lte;
            }
          }
          if (lane < R) d.c_req[(size_t)cc * R + lane] = ev.q;
          if (lane < ITW) d.c_its[(size_t)cc * ITW + lane] = ev.its;
          if (lane == 0) {
            d.c_npods[cc]++;
            d.cver[cc]++;
            set_ord(cpos, cc, cnt(cpos) + 1);
            d.pod_target[pod] = KP_TARGET_CLAIM(cc);
            d.pod_error[pod] = KP_PODERR_NONE;
            sh.pert_kind = PERT_INC;
            sh.pert_pos = cpos;
            sh.found = 1;
            d.counters[1] += cpos + 1;  // claims 0..cpos were evaluated by the reference
          }
          topo_record(d, px, ev.F, d.tmpl_taintset[claim_tmpl(cc)], E + cc, true, lane);
        }
        __syncthreads();
        if (winner >= 0) break;
      }
      __syncthreads();  // the next chunk rewrites warp_mask
    }
    if (tid == 0) {
      if (sh.found)
        commits++;
      else
        ev_inflight += nC;
    }
    if (sh.found) {
      t_inflight += clock64() - t_mark;
      break;
    }

    // ================= addToNewNodeClaim (scheduler.go:592-684) =================
    t_inflight += clock64() - t_mark;
    t_mark = clock64();
    int err = sh.alive_tmpl ? KP_PODERR_INCOMPATIBLE : KP_PODERR_NO_TEMPLATES;
    for (int n = 0; n < d.N && !sh.found; n++) {
      Eval ev;
      ev.ok = false;
      if (warp == 0) {
        uint64_t tw = lane < ITW ? d.tmpl_its[(size_t)n * ITW + lane] : 0ull;
        bool alive = __any_sync(FULL, tw != 0);  // NewScheduler drops templates whose prefilter is empty
        bool skip = !alive;
        uint32_t lp = d.tmpl_limit_present[n];
        if (alive && lp) {  // limits: scheduler.go:605-623, filterByRemainingResources :860-876
          if (d.nodes_res >= 0 && (lp >> d.nodes_res & 1) && d.tmpl_remaining[(size_t)n * R + d.nodes_res] == 0) skip = true;
          if (!skip && lane < ITW) {
            uint64_t keep = 0;
            for (uint64_t bits = tw; bits;) {
              int b = __ffsll((long long)bits) - 1;
              bits &= bits - 1;
              int t = lane * 64 + b;
              bool viable = true;
              for (int r = 0; r < R; r++)
                if ((lp >> r & 1) && d.it_capacity[(size_t)t * R + r] > d.tmpl_remaining[(size_t)n * R + r]) viable = false;
              if (viable) keep |= 1ull << b;
            }
            tw = keep;
          }
          if (!skip && !__any_sync(FULL, tw != 0)) skip = true;
        }
        if (alive && lane == 0) ev_tmpl++;
        if (!skip) {
          int cnew = sh.n_claims;
          if (cnew >= d.Cmax) {
            if (lane == 0) *d.status = KP_ERR_CAPACITY;
          } else if ((px.tmpl_ok >> n) & 1ull) {
            Slot b = lane < K ? rs_slot(d, d.tmpl_rs[n], lane) : slot_absent();
            int64_t bq = lane < R ? d.tmpl_daemon[(size_t)n * R + lane] : 0;
            ev = eval_candidate(d, px, true, b, bq, tw, E + cnew, sh.scratch[0], lane);
          }
        }
        if (lane == 0) sh.ok[0] = ev.ok;
      }
      __syncthreads();
      if (*d.status != KP_OK) {
        if (tid == 0) sh.done = 1;
        __syncthreads();
        break;
      }
      if (sh.ok[0]) {
        const int cnew = sh.n_claims;
        if (warp == 0) {  // NewNodeClaim + Add
          if (lane < K) {
            size_t i = (size_t)cnew * K + lane;
            d.c_sflags[i] = (uint8_t)ev.F.f;
            d.c_smask[i] = ev.F.m;
            if (d.has_bounds) {
              d.c_sgte[i] = ev.F.gte;
              d.c_slte[i] = ev.F.lte;
            }
          }
          if (lane < R) d.c_req[(size_t)cnew * R + lane] = ev.q;
          if (lane < ITW) d.c_its[(size_t)cnew * ITW + lane] = ev.its;
          if (lane == 0) {
            d.c_tmpl[cnew] = n;
            if (cnew < CS) tmpl_s[cnew] = n;
            d.c_npods[cnew] = 1;
            set_ord(cnew, cnew, 1);
            d.pod_target[pod] = KP_TARGET_CLAIM(cnew);
            d.pod_error[pod] = KP_PODERR_NONE;
          }
          // subtractMax (scheduler.go:840-857): remaining -= max capacity over the claim's instance types
          uint32_t lp = d.tmpl_limit_present[n];
          if (lp) {
            for (int r = 0; r < R; r++) {
              if (!(lp >> r & 1)) continue;
              long long mx = 0;
              if (lane < ITW)
                for (uint64_t bits = ev.its; bits;) {
                  int b = __ffsll((long long)bits) - 1;
                  bits &= bits - 1;
                  long long cap = d.it_capacity[(size_t)(lane * 64 + b) * R + r];
                  if (cap > mx) mx = cap;
                }
              for (int o = 16; o; o >>= 1) {
                long long other = __shfl_xor_sync(FULL, mx, o);
                if (other > mx) mx = other;
              }
              if (lane == 0) d.tmpl_remaining[(size_t)n * R + r] -= mx;
            }
          }
        }
        // Topology.Register(hostname) (nodeclaim.go:213): every hostname group learns the new, empty domain
        if (d.GH > 0) {
          for (int g = tid; g < d.G; g += SOLVE_THREADS)
            if (d.groups[g].key == d.hostname_key) {
              d.g_ndomains[g]++;
              d.g_nempty[g]++;
            }
          __syncthreads();
        }
        if (warp == 0) topo_record(d, px, ev.F, d.tmpl_taintset[n], E + cnew, true, lane);
        if (tid == 0) {
          sh.n_claims = cnew + 1;
          sh.pert_kind = PERT_APPEND;
          sh.pert_pos = cnew;
          sh.found = 1;
          commits++;
        }
        __syncthreads();
      }
      __syncthreads();  // sh.ok[0] is rewritten by the next template
    }
    if (sh.done) break;
    if (!sh.found && tid == 0) {  // scheduler.go:415-421: record the error and requeue
      d.pod_error[pod] = (uint8_t)err;
      d.pod_target[pod] = KP_TARGET_UNSCHEDULED;
      d.queue[sh.tail % sh.cap] = pod;
      d.qcls[sh.tail % sh.cap] = px.cls;
      sh.tail++;
      d.last_len[pod] = sh.tail - sh.head;
    }
    } while (0);
    if (sh.done) break;
    if (b_idx >= 0) {  // last warp only: park the staged class row for the next iteration
      store_class_regs(d, sh.ctx[b_idx & 1], br, lane);
      if (lane == 0) sh.ctx_idx[b_idx & 1] = b_idx;
    }
  }
  if (tid == 0) {
    *d.n_claims = sh.n_claims;
    d.counters[0] += ev_existing;
    d.counters[1] += ev_inflight;
    d.counters[2] += ev_tmpl;
    d.counters[3] += commits;
    d.counters[4] += slow_sorts;
    d.counters[5] = t_pop;
    d.counters[6] = t_sort;
    d.counters[7] = t_inflight;
    d.counters[8] = t_new + (clock64() - t_mark) * 0;
    d.counters[9] = n_rounds;
    d.counters[10] = n_ctx_miss;
    d.counters[11] = watchdog;
  }
}
