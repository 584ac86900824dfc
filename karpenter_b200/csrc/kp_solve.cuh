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

struct SolveShared {
  int head, tail, cap;     // circular pod queue
  int pod, cls, done;
  int n_claims;
  int pert_kind, pert_pos; // the one out-of-order element left behind by the previous commit
  int rot_from, rot_to, rot_mode;
  int warp_cnt[SOLVE_WARPS];
  int cand_n;
  int cand[SOLVE_THREADS];     // candidates of the current chunk, in scan order
  int ok[SOLVE_WARPS];
  int winner;
  int found;
  int active_before;           // existing-node evaluation counter helper
  Slot scratch[SOLVE_WARPS][KP_MAXK];
};

// block-wide ordered compaction of a predicate into sh.cand (returns count via sh.cand_n)
__device__ __forceinline__ void compact(SolveShared& sh, bool pass, int value) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned m = __ballot_sync(FULL, pass);
  if (lane == 0) sh.warp_cnt[w] = __popc(m);
  __syncthreads();
  int base = 0, total = 0;
  for (int i = 0; i < SOLVE_WARPS; i++) {
    int c = sh.warp_cnt[i];
    if (i < w) base += c;
    total += c;
  }
  if (pass) sh.cand[base + __popc(m & ((1u << lane) - 1))] = value;
  if (threadIdx.x == 0) sh.cand_n = total;
  __syncthreads();
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(KpDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SolveShared& sh = *reinterpret_cast<SolveShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = d.K, R = d.R, ITW = d.ITW, E = d.E;
  int* order = d.order;
  int* cnt_at = d.cnt_at;
  if (tid == 0) {
    sh.head = 0;
    sh.tail = (int)d.P;
    sh.cap = (int)d.P + 1;
    sh.n_claims = 0;
    sh.pert_kind = PERT_NONE;
    sh.done = 0;
  }
  __syncthreads();
  long long ev_existing = 0, ev_inflight = 0, ev_tmpl = 0, commits = 0, slow_sorts = 0;  // thread 0 only
  int n_active_nodes = 0;
  for (int n = tid; n < E; n += SOLVE_THREADS) n_active_nodes += (d.node_flags[n] & KP_NODE_SCHEDULABLE) ? 1 : 0;
  // (block reduce once)
  {
    for (int o = 16; o; o >>= 1) n_active_nodes += __shfl_xor_sync(FULL, n_active_nodes, o);
    if (lane == 0) sh.warp_cnt[warp] = n_active_nodes;
    __syncthreads();
    n_active_nodes = 0;
    for (int i = 0; i < SOLVE_WARPS; i++) n_active_nodes += sh.warp_cnt[i];
    __syncthreads();
  }

  for (;;) {
    __syncthreads();  // everyone has read sh.found / sh.done of the previous pod
    // ---- Queue.Pop (queue.go:46-60)
    if (tid == 0) {
      int len = sh.tail - sh.head;
      if (len == 0) {
        sh.done = 1;
      } else {
        int pod = d.queue[sh.head % sh.cap];
        if (d.last_len[pod] == len) {
          sh.done = 1;
        } else {
          sh.head++;
          sh.pod = pod;
          sh.cls = d.pod_class[pod];
        }
      }
      sh.found = 0;
    }
    __syncthreads();
    if (sh.done) break;
    const int X = sh.cls, pod = sh.pod;
    const int tolset = d.cls_tolset[X];
    const int rv = d.cls_rv[X];

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
            if (d.cls_req[(size_t)X * R + r] > (present ? rem : 0)) pass = false;
          }
        }
      }
      compact(sh, pass, n);
      int ncand = sh.cand_n;
      for (int g0 = 0; g0 < ncand && !sh.found; g0 += SOLVE_WARPS) {
        int ci = g0 + warp;
        Eval ev;
        ev.ok = false;
        int node = -1;
        if (ci < ncand) {
          node = sh.cand[ci];
          Slot b = lane < K ? load_slot(d.node_sflags, d.node_smask, d.node_sgte, d.node_slte, (size_t)node * K + lane,
                                         d.has_bounds)
                            : slot_absent();
          ev = eval_candidate(d, X, false, b, 0, 0, node, sh.scratch[warp], lane);
        }
        if (lane == 0) sh.ok[warp] = ev.ok;
        __syncthreads();
        if (tid == 0) {
          int w = -1;
          for (int i = 0; i < SOLVE_WARPS && w < 0; i++)
            if (sh.ok[i]) w = i;
          sh.winner = w;
          if (w >= 0) sh.found = 1;
        }
        __syncthreads();
        if (sh.winner == warp) {  // ExistingNode.Add (existingnode.go:147-155)
          if (lane < K) {
            size_t i = (size_t)node * K + lane;
            d.node_sflags[i] = (uint8_t)ev.F.f;
            d.node_smask[i] = ev.F.m;
            if (d.has_bounds) {
              d.node_sgte[i] = ev.F.gte;
              d.node_slte[i] = ev.F.lte;
            }
          }
          if (lane < R) d.node_rem[(size_t)node * R + lane] -= d.cls_req[(size_t)X * R + lane];
          if (lane == 0) {
            d.node_rem_present[node] |= (1u << R) - 1;
            d.node_npods[node]++;
            d.pod_target[pod] = node;
            d.pod_error[pod] = KP_PODERR_NONE;
            d.counters[0] += node + 1;  // nodes 0..node were evaluated by the reference
          }
          topo_record(d, X, ev.F, d.node_taintset[node], node, false, lane);
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      // the reference evaluates every existing node up to the winner (or all of them)
      ev_existing += sh.found ? 0 : n_active_nodes;  // (winner-prefix counts are added by the host from pod_target)
      if (sh.found) commits++;
    }
    if (sh.found) continue;

    // ================= sort.Slice(newNodeClaims, len(Pods) asc) (scheduler.go:504) =================
    const int nC = sh.n_claims;
    if (sh.pert_kind != PERT_NONE) {
      if (tid == 0) {
        int p = sh.pert_pos;
        bool inversion = sh.pert_kind == PERT_INC ? (p + 1 < nC && cnt_at[p + 1] < cnt_at[p])
                                                  : (nC >= 2 && cnt_at[nC - 1] < cnt_at[nC - 2]);
        sh.rot_mode = 0;
        if (inversion) {
          bool stable = d.stable_order || nC <= 12;
          if (!stable && nC >= 50) {
            DevSorter s{cnt_at, order};
            int hint;
            s.choose_pivot(0, nC, &hint);
            stable = hint == 1;  // partialInsertionSort repairs a single inversion == stable move
          }
          if (stable) {
            if (sh.pert_kind == PERT_INC) {
              int c = cnt_at[p];  // elevated count; move right past every smaller element
              int lo = p + 1, hi = nC;
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cnt_at[mid] < c)
                  lo = mid + 1;
                else
                  hi = mid;
              }
              sh.rot_from = p;
              sh.rot_to = lo - 1;  // new position of the elevated element
              sh.rot_mode = 1;
            } else {
              int c = cnt_at[nC - 1];  // new claim: move left past every larger element
              int lo = 0, hi = nC - 1;
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cnt_at[mid] <= c)
                  lo = mid + 1;
                else
                  hi = mid;
              }
              sh.rot_from = nC - 1;
              sh.rot_to = lo;
              sh.rot_mode = 2;
            }
          } else {
            DevSorter s{cnt_at, order};
            s.pdqsort(0, nC, DevSorter::bits_len((unsigned long long)nC));
            slow_sorts++;
          }
        }
        sh.pert_kind = PERT_NONE;
      }
      __syncthreads();
      if (sh.rot_mode == 1) {  // rotate [from, to] left by one
        int from = sh.rot_from, to = sh.rot_to;
        int eo = order[from], ec = cnt_at[from];
        for (int b0 = from; b0 < to; b0 += SOLVE_THREADS) {
          int i = b0 + tid;
          int vo = 0, vc = 0;
          if (i < to) {
            vo = order[i + 1];
            vc = cnt_at[i + 1];
          }
          __syncthreads();
          if (i < to) {
            order[i] = vo;
            cnt_at[i] = vc;
          }
          __syncthreads();
        }
        if (tid == 0) {
          order[to] = eo;
          cnt_at[to] = ec;
        }
        __syncthreads();
      } else if (sh.rot_mode == 2) {  // rotate [to, from] right by one
        int from = sh.rot_from, to = sh.rot_to;
        int eo = order[from], ec = cnt_at[from];
        for (int b0 = from; b0 > to; b0 -= SOLVE_THREADS) {
          int i = b0 - tid;
          int vo = 0, vc = 0;
          if (i > to) {
            vo = order[i - 1];
            vc = cnt_at[i - 1];
          }
          __syncthreads();
          if (i > to) {
            order[i] = vo;
            cnt_at[i] = vc;
          }
          __syncthreads();
        }
        if (tid == 0) {
          order[to] = eo;
          cnt_at[to] = ec;
        }
        __syncthreads();
      }
    }

    // ================= addToInflightNode (scheduler.go:557-589) =================
    const int rdw = (d.Cmax + 31) >> 5;
    for (int base = 0; base < nC && !sh.found; base += SOLVE_THREADS) {
      int pos = base + tid;
      bool pass = false;
      int c = -1;
      if (pos < nC) {
        c = order[pos];
        pass = !(d.rdead[(size_t)rv * rdw + (c >> 5)] >> (c & 31) & 1);
        if (pass) pass = tolerated(d, tolset, d.tmpl_taintset[d.c_tmpl[c]]);
      }
      compact(sh, pass, pos);
      int ncand = sh.cand_n;
      for (int g0 = 0; g0 < ncand && !sh.found; g0 += SOLVE_WARPS) {
        int ci = g0 + warp;
        Eval ev;
        ev.ok = false;
        ev.res_dead = false;
        int cpos = -1, cc = -1;
        if (ci < ncand) {
          cpos = sh.cand[ci];
          cc = order[cpos];
          Slot b = lane < K ? load_slot(d.c_sflags, d.c_smask, d.c_sgte, d.c_slte, (size_t)cc * K + lane, d.has_bounds)
                            : slot_absent();
          int64_t bq = lane < R ? d.c_req[(size_t)cc * R + lane] : 0;
          uint64_t bi = lane < ITW ? d.c_its[(size_t)cc * ITW + lane] : 0ull;
          ev = eval_candidate(d, X, true, b, bq, bi, E + cc, sh.scratch[warp], lane);
          if (ev.res_dead && lane == 0) atomicOr(&d.rdead[(size_t)rv * rdw + (cc >> 5)], 1u << (cc & 31));
        }
        if (lane == 0) sh.ok[warp] = ev.ok;
        __syncthreads();
        if (tid == 0) {
          int w = -1;
          for (int i = 0; i < SOLVE_WARPS && w < 0; i++)
            if (sh.ok[i]) w = i;
          sh.winner = w;
          if (w >= 0) sh.found = 1;
        }
        __syncthreads();
        if (sh.winner == warp) {  // NodeClaim.Add (nodeclaim.go:207-219)
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
            cnt_at[cpos]++;
            d.pod_target[pod] = KP_TARGET_CLAIM(cc);
            d.pod_error[pod] = KP_PODERR_NONE;
            sh.pert_kind = PERT_INC;
            sh.pert_pos = cpos;
          }
          topo_record(d, X, ev.F, d.tmpl_taintset[d.c_tmpl[cc]], E + cc, true, lane);
          if (lane == 0) d.counters[1] += cpos + 1;  // claims 0..cpos were evaluated by the reference
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      if (sh.found)
        commits++;
      else
        ev_inflight += nC;
    }
    if (sh.found) continue;

    // ================= addToNewNodeClaim (scheduler.go:592-684) =================
    int err = KP_PODERR_NO_TEMPLATES;
    for (int n = 0; n < d.N && !sh.found; n++) {
      // template alive? (NewScheduler drops templates whose prefilter is empty, scheduler.go:148-157)
      uint64_t tw = (warp == 0 && lane < ITW) ? d.tmpl_its[(size_t)n * ITW + lane] : 0ull;
      bool alive = false;
      Eval ev;
      ev.ok = false;
      if (warp == 0) {
        alive = __any_sync(FULL, tw != 0);
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
        if (lane == 0) sh.ok[1] = alive;
        if (!skip) {
          int cnew = sh.n_claims;
          if (cnew >= d.Cmax) {
            if (lane == 0) *d.status = KP_ERR_CAPACITY;
          } else if (tolerated(d, tolset, d.tmpl_taintset[n])) {
            Slot b = lane < K ? rs_slot(d, d.tmpl_rs[n], lane) : slot_absent();
            int64_t bq = lane < R ? d.tmpl_daemon[(size_t)n * R + lane] : 0;
            ev = eval_candidate(d, X, true, b, bq, tw, E + cnew, sh.scratch[0], lane);
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
      if (sh.ok[1]) {
        err = KP_PODERR_INCOMPATIBLE;
        if (tid == 0) ev_tmpl++;
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
            d.c_npods[cnew] = 1;
            order[cnew] = cnew;
            cnt_at[cnew] = 1;
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
        for (int g = tid; g < d.G; g += SOLVE_THREADS)
          if (d.groups[g].key == d.hostname_key) {
            d.g_ndomains[g]++;
            d.g_nempty[g]++;
          }
        __syncthreads();
        if (warp == 0) topo_record(d, X, ev.F, d.tmpl_taintset[n], E + cnew, true, lane);
        if (tid == 0) {
          sh.n_claims = cnew + 1;
          sh.pert_kind = PERT_APPEND;
          sh.pert_pos = cnew;
          sh.found = 1;
          commits++;
        }
        __syncthreads();
      }
    }
    if (sh.done) break;
    if (!sh.found && tid == 0) {  // scheduler.go:415-421: record the error and requeue
      d.pod_error[pod] = (uint8_t)err;
      d.pod_target[pod] = KP_TARGET_UNSCHEDULED;
      d.queue[sh.tail % sh.cap] = pod;
      sh.tail++;
      d.last_len[pod] = sh.tail - sh.head;
    }
    __syncthreads();
  }
  if (tid == 0) {
    *d.n_claims = sh.n_claims;
    d.counters[0] += ev_existing;
    d.counters[1] += ev_inflight;
    d.counters[2] += ev_tmpl;
    d.counters[3] += commits;
    d.counters[4] += slow_sorts;
  }
}
