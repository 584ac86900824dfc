// TEST INFRASTRUCTURE -- CPU oracle. Not part of the product: only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may build, load or call this library (liborc.so).
//
// orc_solver.cpp: restatement of the reference's provisioning hot path on the flat kp_problem input:
//   Scheduler.Solve / add / addToExistingNode / addToInflightNode / addToNewNodeClaim
//                                 pkg/controllers/provisioning/scheduling/scheduler.go:381-684
//   NodeClaim.CanAdd / Add / filterInstanceTypesByRequirements      .../nodeclaim.go:83-219,412-488
//   ExistingNode.CanAdd / Add                                        .../existingnode.go:40-155
//   Queue                                                            .../queue.go:37-108
//   SimulateScheduling + computeConsolidation          pkg/controllers/disruption/helpers.go:51-142,
//                                                      consolidation.go:136-229,319-337
// PARITY PINNING: the requirement algebra is pinned by the reference's own known-answer tables
// (tests/golden/requirement_kats.json, extracted from pkg/scheduling/requirement_test.go and requirements_test.go);
// solver-level behaviour is pinned by the reference's aggregate assertions restated in tests/test_oracle_*.py.
// Go's sort.Slice tie order (orc_gosort.hpp) and Go map iteration order are UNPINNED by construction.
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>

#include "orc_gosort.hpp"
#include "orc_topology.hpp"

namespace orc {

struct InflightClaim {
  int tmpl;
  std::vector<int> reserved;  // reservation ids of NodeClaim.reservedOfferings (ascending)
  uint64_t ports = 0;         // hostPortUsage: interned <ip, port, protocol> entries in use (daemons + pods)
  Requirements reqs;
  std::vector<int> its;  // InstanceTypeOptions (global ids, template order)
  Res requests;          // Spec.Resources.Requests
  std::vector<int64_t> pods;
  int32_t hostname;
  int created;  // creation index
};

struct ExistingNode {
  uint64_t ports = 0;  // StateNode.HostPortUsage()
  int node;
  Res remaining;
  Requirements reqs;
  int taintset;
  std::vector<int64_t> pods;
};

struct Counters {
  int64_t existing = 0, inflight = 0, tmpl = 0, commits = 0;
};

// parallelizeUntil (scheduler.go:757-779): W workers pull candidate indices from a shared counter; the lowest index
// that succeeds wins and every index below it has been evaluated.  Persistent spinning workers, one job at a time.
struct WorkerPool {
  int W;
  std::vector<std::thread> threads;
  std::atomic<uint64_t> epoch{0};
  std::atomic<int> finished{0};
  std::atomic<bool> quit{false};
  std::function<void(int)> job;
  explicit WorkerPool(int w) : W(w) {
    for (int i = 1; i < W; i++)
      threads.emplace_back([this, i] {
        uint64_t seen = 0;
        for (;;) {
          uint64_t e;
          while ((e = epoch.load(std::memory_order_acquire)) == seen && !quit.load(std::memory_order_relaxed)) {
          }
          if (quit.load(std::memory_order_relaxed)) return;
          seen = e;
          job(i);
          finished.fetch_add(1, std::memory_order_release);
        }
      });
  }
  ~WorkerPool() {
    quit.store(true);
    for (auto& t : threads) t.join();
  }
  void run(const std::function<void(int)>& f) {
    job = f;
    finished.store(0, std::memory_order_relaxed);
    epoch.fetch_add(1, std::memory_order_release);
    f(0);
    while (finished.load(std::memory_order_acquire) != W - 1) {
    }
  }
};

struct Scheduler {
  const Prob& P;
  const kp_problem* p;
  Topology topo;
  std::vector<std::vector<int>> tmpl_options;  // NodeClaimTemplate.InstanceTypeOptions after the prefilter
  std::vector<uint8_t> tmpl_alive;
  std::vector<Res> remaining;  // remainingResources per NodePool (limits)
  std::vector<InflightClaim*> new_claims;  // s.newNodeClaims (sorted in place by add())
  std::vector<std::unique_ptr<InflightClaim>> claim_store;
  std::vector<ExistingNode> existing;
  int64_t hostname_seq = 0;
  Counters ctr;
  bool stable_order;
  WorkerPool* pool = nullptr;  // null: evaluate candidates serially
  // ReservationManager (reservationmanager.go:28-110): remaining capacity per reservation id; which ids a NodeClaim
  // holds lives on the claim (InflightClaim::reserved == reservations[hostname])
  std::vector<int> rsv_cap;
  bool rsv_strict;

  Scheduler(const Prob& prob) : P(prob), p(prob.p), topo(prob), stable_order(prob.p->claim_order_mode == 1) {
    rsv_strict = p->reserved_offering_strict != 0;
    // NewReservationManager (:33-52): the least capacity any offering of the id reports
    rsv_cap.assign(p->n_reservations > 0 ? p->n_reservations : 0, -1);
    if (p->off_reserved && p->off_reservation_id)
      for (int t = 0; t < p->n_its; t++)
        for (int o = p->it_off_off[t]; o < p->it_off_off[t + 1]; o++) {
          if (!p->off_reserved[o]) continue;
          const int id = p->off_reservation_id[o], cap = p->off_reservation_capacity[o];
          if (rsv_cap[id] < 0 || rsv_cap[id] > cap) rsv_cap[id] = cap;
        }
    for (int& c : rsv_cap)
      if (c < 0) c = 0;
  }

  // HostPortUsage.Conflicts (hostportusage.go:75-88): some port of the pod Matches a port in use
  uint64_t class_ports(int cls) const { return p->class_hostports ? p->class_hostports[cls] : 0ull; }
  bool ports_conflict(uint64_t used, int cls) const {
    uint64_t mine = class_ports(cls);
    while (mine) {
      const int i = __builtin_ctzll(mine);
      mine &= mine - 1;
      if (used & p->hostport_conflicts[i]) return true;
    }
    return false;
  }

  // offeringsToReserve (nodeclaim.go:240-287).  false == ReservedOfferingError (strict mode only)
  bool offerings_to_reserve(const InflightClaim& c, const std::vector<int>& its, const Requirements& reqs,
                            std::vector<int>* out) const {
    out->clear();
    if (rsv_cap.empty()) return true;
    bool has_compatible = false;
    std::set<int> ids;
    for (int it : its)
      for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++) {
        if (!p->off_reserved[o] || !p->off_available[o]) continue;
        if (!compatible(P, P, reqs, P.reqsets[p->off_reqset[o]], true)) continue;
        has_compatible = true;
        const int id = p->off_reservation_id[o];
        const bool held = std::find(c.reserved.begin(), c.reserved.end(), id) != c.reserved.end();
        if (held || rsv_cap[id] > 0) ids.insert(id);  // ReservationManager.CanReserve (:55-70)
      }
    if (rsv_strict) {
      if (has_compatible && ids.empty()) return false;
      if (!c.reserved.empty() && ids.empty()) return false;
    }
    out->assign(ids.begin(), ids.end());
    return true;
  }

  bool offering_compatible(const Requirements& reqs, int it) const {
    for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++)
      if (p->off_available[o] && compatible(P, P, reqs, P.reqsets[p->off_reqset[o]], true)) return true;
    return false;
  }

  // InstanceTypes.SatisfiesMinValues (cloudprovider/types.go:301-337): for every requirement that carries minValues, the
  // instance types together must offer at least that many distinct values of its key
  bool satisfies_min_values(const std::vector<int>& its, const Requirements& reqs) const {
    for (auto& kv : reqs.m) {
      if (!kv.second.has_min) continue;
      int m = -1;
      for (int i = 0; i < p->n_minvalue_keys; i++)
        if (p->minvalue_key[i] == kv.first) m = i;
      std::set<int32_t> values;
      if (m >= 0)
        for (int it : its) {
          size_t row = (size_t)m * p->n_its + it;
          values.insert(p->minvalue_it_vals + p->minvalue_it_off[row], p->minvalue_it_vals + p->minvalue_it_off[row + 1]);
        }
      if ((int)values.size() < kv.second.min_values) return false;
    }
    return true;
  }

  // nodeclaim.go:412-480
  std::vector<int> filter_instance_types(const std::vector<int>& its, const Requirements& reqs, const Res& total) const {
    std::vector<int> remaining;
    for (int it : its) {
      bool it_compat = intersects(P, P.reqsets[p->it_reqset[it]], reqs);  // nodeclaim.go:482-484
      bool it_fits = fits(P.R, total, P.it_alloc[it]);                    // nodeclaim.go:486-488
      bool it_off = offering_compatible(reqs, it);
      if (it_compat && it_fits && it_off) remaining.push_back(it);
    }
    // nodeclaim.go:464-475: Strict refuses the NodeClaim, BestEffort lowers minValues and carries on
    if (!p->min_values_best_effort && reqs.has_min_values() && !satisfies_min_values(remaining, reqs)) remaining.clear();
    return remaining;
  }

  // scheduler.go:116-184 NewScheduler
  void init(const std::vector<uint8_t>& node_active, const std::vector<std::pair<int, int>>& bound_pods,
            const std::vector<int>& pending_classes) {
    int N = p->n_templates;
    tmpl_options.resize(N);
    tmpl_alive.assign(N, 0);
    remaining.resize(N);
    Res zero;
    for (int n = 0; n < N; n++) {
      std::vector<int> its(p->tmpl_its + p->tmpl_it_off[n], p->tmpl_its + p->tmpl_it_off[n + 1]);
      tmpl_options[n] = filter_instance_types(its, P.reqsets[p->tmpl_reqset[n]], zero);  // scheduler.go:147
      tmpl_alive[n] = !tmpl_options[n].empty();
      remaining[n] = P.res_row(p->tmpl_limits, n, p->tmpl_limit_present ? p->tmpl_limit_present[n] : 0);
    }
    topo.bound_pods = bound_pods;
    topo.state_node = node_active;
    topo.init(pending_classes);
    // calculateExistingNodeClaims (scheduler.go:686-695); input order == sortExistingNodes order
    for (int i = 0; i < p->n_nodes; i++) {
      if (!node_active[i]) continue;
      ExistingNode e;
      e.node = i;
      e.remaining = P.res_row(p->node_available, i, p->node_avail_present ? p->node_avail_present[i] : P.all_mask());
      e.reqs = topo.node_label_reqs(i);  // existingnode.go:61-63
      e.taintset = p->node_taintset[i];
      e.ports = p->node_hostports ? p->node_hostports[i] : 0ull;
      if (P.hostname_key >= 0) topo.register_domain(P.hostname_key, p->node_hostname[i]);
      existing.push_back(std::move(e));
      int t = p->node_template ? p->node_template[i] : -1;  // updateRemainingResources (scheduler.go:728-735)
      if (t >= 0 && p->node_capacity)
        for (int r = 0; r < P.R; r++)
          if (remaining[t].present >> r & 1) remaining[t].v[r] -= p->node_capacity[(int64_t)i * P.R + r];
    }
  }

  Res class_requests(int cls) const { return P.res_row(p->class_requests, cls, P.all_mask()); }

  // existingnode.go:70-143
  // volume-topology alternatives (nodeclaim.go:136-153, existingnode.go:98-113): class_vol_next chains the classes that carry
  // the same pod with the next alternative folded into class_reqset; a candidate that rejects one is offered the next
  int vol_next(int x) const { return p->class_vol_next ? p->class_vol_next[x] : -1; }
  bool existing_can_add(ExistingNode& n, int cls, Requirements* out) {
    ctr.existing++;
    for (int x = cls; x >= 0; x = vol_next(x))
      if (existing_can_add_alt(n, cls, x, out)) return true;
    return false;
  }
  // cls: the pod's class (tolerations, ports, requests, topology); x: the class whose requirement set is tried
  bool existing_can_add_alt(ExistingNode& n, int cls, int x, Requirements* out) {
    if (!P.tolerates(n.taintset, p->class_tolset[cls])) return false;
    if (ports_conflict(n.ports, cls)) return false;  // existingnode.go:76-82
    if (!fits(P.R, class_requests(cls), n.remaining)) return false;
    const Requirements& pod_reqs = P.reqsets[p->class_reqset[x]];
    if (!compatible(P, P, n.reqs, pod_reqs, false)) return false;
    Requirements node_reqs = n.reqs;
    node_reqs.add_all(P, pod_reqs);
    Requirements topo_reqs;
    if (!topo.add_requirements(cls, n.taintset, P.reqsets[p->class_strict_reqset[cls]], node_reqs, &topo_reqs))
      return false;
    if (!compatible(P, P, node_reqs, topo_reqs, false)) return false;
    node_reqs.add_all(P, topo_reqs);
    *out = node_reqs;
    return true;
  }

  // nodeclaim.go:114-202
  bool claim_can_add(InflightClaim& c, int cls, Requirements* out_reqs, std::vector<int>* out_its,
                     std::vector<int>* out_rsv = nullptr, bool* rsv_error = nullptr) {
    for (int x = cls; x >= 0; x = vol_next(x))  // the LAST alternative's error is the one CanAdd returns (nodeclaim.go:145-152)
      if (claim_can_add_alt(c, cls, x, out_reqs, out_its, out_rsv, rsv_error)) return true;
    return false;
  }
  bool claim_can_add_alt(InflightClaim& c, int cls, int x, Requirements* out_reqs, std::vector<int>* out_its,
                         std::vector<int>* out_rsv, bool* rsv_error) {
    if (rsv_error) *rsv_error = false;
    int taintset = p->tmpl_taintset[c.tmpl];
    if (!P.tolerates(taintset, p->class_tolset[cls])) return false;
    if (ports_conflict(c.ports, cls)) return false;  // nodeclaim.go:120-124
    const Requirements& pod_reqs = P.reqsets[p->class_reqset[x]];
    Requirements reqs = c.reqs;
    if (!compatible(P, P, reqs, pod_reqs, true)) return false;
    reqs.add_all(P, pod_reqs);
    Requirements topo_reqs;
    if (!topo.add_requirements(cls, taintset, P.reqsets[p->class_strict_reqset[cls]], reqs, &topo_reqs)) return false;
    if (!compatible(P, P, reqs, topo_reqs, true)) return false;
    reqs.add_all(P, topo_reqs);
    Res total = merge(P.R, c.requests, class_requests(cls));
    std::vector<int> rem = filter_instance_types(c.its, reqs, total);
    if (rem.empty()) return false;
    std::vector<int> rsv;
    if (!offerings_to_reserve(c, rem, reqs, &rsv)) {  // nodeclaim.go:197-200
      if (rsv_error) *rsv_error = true;
      return false;
    }
    if (out_rsv) *out_rsv = rsv;
    *out_reqs = reqs;
    *out_its = rem;
    return true;
  }

  // nodeclaim.go:207-219
  void claim_add(InflightClaim& c, int64_t pod, int cls, const Requirements& reqs, const std::vector<int>& its,
                 const std::vector<int>& rsv = {}) {
    // Reserve the new ids, release the ones the tighter requirements ruled out (nodeclaim.go:216-218, :223-235)
    for (int id : rsv)
      if (std::find(c.reserved.begin(), c.reserved.end(), id) == c.reserved.end()) rsv_cap[id]--;
    for (int id : c.reserved)
      if (std::find(rsv.begin(), rsv.end(), id) == rsv.end()) rsv_cap[id]++;
    c.reserved = rsv;
    c.ports |= class_ports(cls);  // hostPortUsage.Add (nodeclaim.go:215)
    c.pods.push_back(pod);
    c.its = its;
    c.requests = merge(P.R, c.requests, class_requests(cls));
    c.reqs = reqs;
    if (P.hostname_key >= 0) topo.register_domain(P.hostname_key, c.hostname);
    topo.record(cls, p->tmpl_taintset[c.tmpl], reqs);
    ctr.commits++;
  }

  // scheduler.go:520-555
  bool add_to_existing(int64_t pod, int cls, int* target) {
    for (size_t i = 0; i < existing.size(); i++) {
      Requirements r;
      if (existing_can_add(existing[i], cls, &r)) {
        ExistingNode& n = existing[i];
        n.pods.push_back(pod);
        Res q = class_requests(cls);
        for (int k = 0; k < P.R; k++) n.remaining.v[k] -= q.v[k];  // SubtractFrom creates missing keys
        n.remaining.present |= P.all_mask();
        n.ports |= class_ports(cls);  // existingnode.go:153
        n.reqs = r;
        topo.record(cls, n.taintset, r);
        ctr.commits++;
        *target = n.node;
        return true;
      }
    }
    return false;
  }

  void sort_claims() {
    auto less = [&](int a, int b) { return new_claims[a]->pods.size() < new_claims[b]->pods.size(); };
    if (stable_order) {
      std::stable_sort(new_claims.begin(), new_claims.end(),
                       [](InflightClaim* a, InflightClaim* b) { return a->pods.size() < b->pods.size(); });
    } else {
      go_sort_slice((int)new_claims.size(), less, [&](int a, int b) { std::swap(new_claims[a], new_claims[b]); });
    }
  }

  // scheduler.go:557-589 with parallelizeUntil: same result as the serial scan (lowest succeeding index)
  bool add_to_inflight_parallel(int64_t pod, int cls, int* target) {
    const int n = (int)new_claims.size();
    std::atomic<int> next{0}, best{INT32_MAX};
    struct Slot_ {
      int idx = INT32_MAX;
      Requirements r;
      std::vector<int> its, rsv;
    };
    std::vector<Slot_> slots(pool->W);
    pool->run([&](int w) {
      for (;;) {
        int i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= n || i > best.load(std::memory_order_relaxed)) return;
        Requirements r;
        std::vector<int> its, rsv;
        if (claim_can_add(*new_claims[i], cls, &r, &its, &rsv)) {
          if (i < slots[w].idx) {
            slots[w].idx = i;
            slots[w].r = std::move(r);
            slots[w].its = std::move(its);
            slots[w].rsv = std::move(rsv);
          }
          int cur = best.load();
          while (i < cur && !best.compare_exchange_weak(cur, i)) {
          }
          return;
        }
      }
    });
    int b = best.load();
    if (b == INT32_MAX) {
      ctr.inflight += n;
      return false;
    }
    ctr.inflight += b + 1;
    for (auto& sl : slots)
      if (sl.idx == b) {
        claim_add(*new_claims[b], pod, cls, sl.r, sl.its, sl.rsv);
        *target = KP_TARGET_CLAIM(new_claims[b]->created);
        return true;
      }
    return false;
  }

  // scheduler.go:557-589
  bool add_to_inflight(int64_t pod, int cls, int* target) {
    if (pool && new_claims.size() >= 32) return add_to_inflight_parallel(pod, cls, target);
    for (size_t i = 0; i < new_claims.size(); i++) {
      Requirements r;
      std::vector<int> its, rsv;
      ctr.inflight++;
      if (claim_can_add(*new_claims[i], cls, &r, &its, &rsv)) {
        claim_add(*new_claims[i], pod, cls, r, its, rsv);
        *target = KP_TARGET_CLAIM(new_claims[i]->created);
        return true;
      }
    }
    return false;
  }

  // scheduler.go:592-684
  bool add_to_new_claim(int64_t pod, int cls, int* target, bool* reserved_error) {
    *reserved_error = false;
    for (int n = 0; n < p->n_templates; n++) {
      if (!tmpl_alive[n]) continue;
      ctr.tmpl++;
      std::vector<int> its = tmpl_options[n];
      Res& rem = remaining[n];
      if (P.nodes_res >= 0 && (rem.present >> P.nodes_res & 1) && rem.v[P.nodes_res] == 0) continue;  // :607-611
      if (rem.present) {  // filterByRemainingResources (:860-876)
        std::vector<int> f;
        for (int it : its) {
          bool viable = true;
          for (int r = 0; r < P.R; r++)
            if ((rem.present >> r & 1)) {
              uint32_t cp = p->it_cap_present ? p->it_cap_present[it] : P.all_mask();
              int64_t cap = (cp >> r & 1) ? p->it_capacity[(int64_t)it * P.R + r] : 0;
              if (cap > rem.v[r]) viable = false;
            }
          if (viable) f.push_back(it);
        }
        its.swap(f);
        if (its.empty()) continue;
      }
      // NewNodeClaim (nodeclaim.go:83-109)
      InflightClaim c;
      c.tmpl = n;
      c.reqs = P.reqsets[p->tmpl_reqset[n]];
      c.hostname = (P.hostname_key >= 0 ? P.nvalues(P.hostname_key) : 0) + (int32_t)(hostname_seq++);
      if (P.hostname_key >= 0) {
        Requirement h;
        h.key = P.hostname_key;
        h.values.push_back(c.hostname);
        c.reqs.add(P, h);
      }
      c.its = its;
      c.ports = p->tmpl_hostports ? p->tmpl_hostports[n] : 0ull;  // daemonHostPortUsage (scheduler.go:794-811)
      c.requests = P.res_row(p->tmpl_daemon, n, P.all_mask());
      Requirements r;
      std::vector<int> rem_its, rsv;
      bool rsv_err = false;
      if (!claim_can_add(c, cls, &r, &rem_its, &rsv, &rsv_err)) {
        // A NodePool with compatible reserved capacity that is taken: no fallback to a NodePool of lower weight
        // (scheduler.go:632-646) -- the lowest index that succeeded OR reported this error decides
        if (rsv_err) {
          *reserved_error = true;
          return false;
        }
        continue;
      }
      auto owned = std::make_unique<InflightClaim>(std::move(c));
      owned->created = (int)claim_store.size();
      claim_add(*owned, pod, cls, r, rem_its, rsv);
      new_claims.push_back(owned.get());
      *target = KP_TARGET_CLAIM(owned->created);
      // subtractMax (scheduler.go:840-857)
      for (int k = 0; k < P.R; k++) {
        if (!(rem.present >> k & 1)) continue;
        int64_t mx = 0;
        bool any = false;
        for (int it : owned->its) {
          uint32_t cp = p->it_cap_present ? p->it_cap_present[it] : P.all_mask();
          if (!(cp >> k & 1)) continue;
          int64_t cap = p->it_capacity[(int64_t)it * P.R + k];
          if (!any || cap > mx) mx = cap;
          any = true;
        }
        rem.v[k] -= mx;
      }
      claim_store.push_back(std::move(owned));
      return true;
    }
    return false;
  }

  // scheduler.go:493-518
  bool add(int64_t pod, int cls, int* target, uint8_t* err) {
    if (add_to_existing(pod, cls, target)) return true;
    sort_claims();  // scheduler.go:504
    if (add_to_inflight(pod, cls, target)) return true;
    bool any_template = false;
    for (int n = 0; n < p->n_templates; n++) any_template |= (bool)tmpl_alive[n];
    if (!any_template) {
      *err = KP_PODERR_NO_TEMPLATES;
      return false;
    }
    bool reserved_error = false;
    if (add_to_new_claim(pod, cls, target, &reserved_error)) return true;
    *err = reserved_error ? KP_PODERR_RESERVED : KP_PODERR_INCOMPATIBLE;
    return false;
  }

  // queue.go:72-108 byCPUAndMemoryDescending
  bool queue_less(int64_t a, int64_t b) const {
    int ca = p->pod_class[a], cb = p->pod_class[b];
    int64_t cpu_a = P.cpu_res >= 0 ? p->class_requests[(int64_t)ca * P.R + P.cpu_res] : 0;
    int64_t cpu_b = P.cpu_res >= 0 ? p->class_requests[(int64_t)cb * P.R + P.cpu_res] : 0;
    if (cpu_a != cpu_b) return cpu_a > cpu_b;
    int64_t mem_a = P.mem_res >= 0 ? p->class_requests[(int64_t)ca * P.R + P.mem_res] : 0;
    int64_t mem_b = P.mem_res >= 0 ? p->class_requests[(int64_t)cb * P.R + P.mem_res] : 0;
    if (mem_a != mem_b) return mem_a > mem_b;
    int64_t ta = p->pod_creation ? p->pod_creation[a] : 0, tb = p->pod_creation ? p->pod_creation[b] : 0;
    if (ta != tb) return ta < tb;
    if (p->pod_uid_hi[a] != p->pod_uid_hi[b]) return p->pod_uid_hi[a] < p->pod_uid_hi[b];
    return p->pod_uid_lo[a] < p->pod_uid_lo[b];
  }

  // scheduler.go:381-436 Solve over the given pod rows
  void solve(const std::vector<int64_t>& pod_rows, std::vector<int32_t>* targets, std::vector<uint8_t>* errors) {
    std::vector<int64_t> q = pod_rows;
    std::sort(q.begin(), q.end(), [&](int64_t a, int64_t b) { return queue_less(a, b); });  // total order
    std::unordered_map<int64_t, size_t> last_len;
    std::unordered_map<int64_t, size_t> slot;
    for (size_t i = 0; i < pod_rows.size(); i++) slot[pod_rows[i]] = i;
    targets->assign(pod_rows.size(), KP_TARGET_UNSCHEDULED);
    errors->assign(pod_rows.size(), KP_PODERR_NONE);
    size_t head = 0;
    for (;;) {
      size_t len = q.size() - head;
      if (len == 0) break;
      int64_t pod = q[head];
      auto ll = last_len.find(pod);
      if (ll != last_len.end() && ll->second == len) break;  // queue.go:54-58: a full cycle without progress
      head++;
      const int cls0 = p->pod_class[pod];
      int cls = cls0;
      int target = KP_TARGET_UNSCHEDULED;
      uint8_t err = KP_PODERR_NONE;
      bool placed = false;
      for (;;) {  // trySchedule (scheduler.go:438-469): relax one soft constraint at a time until the pod fits
        if (add(pod, cls, &target, &err)) {
          placed = true;
          break;
        }
        if (err == KP_PODERR_RESERVED) break;  // never relax on a ReservedOfferingError (scheduler.go:447-453)
        int nx = p->class_relax_next ? p->class_relax_next[cls] : -1;  // Preferences.Relax (preferences.go:38-57)
        if (nx < 0) break;
        cls = nx;
        topo.update(cls);
      }
      if (placed) {
        (*targets)[slot[pod]] = target;
        (*errors)[slot[pod]] = KP_PODERR_NONE;
      } else {
        (*errors)[slot[pod]] = err;
        topo.update(cls0);  // the ORIGINAL pod goes back into the queue (scheduler.go:415-421)
        q.push_back(pod);  // queue.go:63-66 Push
        last_len[pod] = q.size() - head;
      }
    }
    // FinalizeScheduling (nodeclaim.go:291-307): drop the hostname requirement; a claim that holds reservations is
    // pinned to capacity-type In [reserved] and reservation-id In [held ids]
    if (P.hostname_key >= 0)
      for (auto* c : new_claims) c->reqs.m.erase(P.hostname_key);
    if (!rsv_cap.empty() && p->reservation_capacity_type_key >= 0 && p->reservation_id_key >= 0 && p->reservation_value)
      for (auto* c : new_claims) {
        if (c->reserved.empty()) continue;
        Requirement ct;
        ct.key = p->reservation_capacity_type_key;
        ct.values.push_back(p->reservation_reserved_value);
        c->reqs.m[ct.key] = ct;  // overwritten, not intersected (:300)
        Requirement rid;
        rid.key = p->reservation_id_key;
        for (int id : c->reserved) rid.values.push_back(p->reservation_value[id]);
        std::sort(rid.values.begin(), rid.values.end());
        c->reqs.add(P, rid);
      }
  }
};

// ---- price helpers (pkg/cloudprovider/types.go) ----
struct Pricing {
  const Prob& P;
  const kp_problem* p;
  explicit Pricing(const Prob& prob) : P(prob), p(prob.p) {}
  // types.go:238-257 OrderByPrice comparator key: cheapest available compatible offering, MaxFloat64 if none
  double min_price(int it, const Requirements& reqs) const {
    double best = 1.7976931348623157e308;
    for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++)
      if (p->off_available[o] && compatible(P, P, reqs, P.reqsets[p->off_reqset[o]], true) && p->off_price[o] < best)
        best = p->off_price[o];
    return best;
  }
  void order_by_price(std::vector<int>& its, const Requirements& reqs) const {
    go_sort_slice((int)its.size(), [&](int a, int b) { return min_price(its[a], reqs) < min_price(its[b], reqs); },
                  [&](int a, int b) { std::swap(its[a], its[b]); });
  }
  // types.go:480-491 WorstLaunchPrice over Offerings.Available()
  double worst_launch_price(int it, const Requirements& reqs, int ct_key, const int ct_order[3]) const {
    for (int i = 0; i < 3; i++) {
      if (ct_key < 0 || ct_order[i] < 0) continue;
      Requirements ct;
      Requirement r;
      r.key = ct_key;
      r.values.push_back(ct_order[i]);
      ct.add(P, r);
      bool any = false;
      double worst = 0;
      for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++) {
        if (!p->off_available[o]) continue;
        const Requirements& oreq = P.reqsets[p->off_reqset[o]];
        if (!compatible(P, P, reqs, oreq, true)) continue;
        if (!compatible(P, P, ct, oreq, true)) continue;
        if (!any || p->off_price[o] > worst) worst = p->off_price[o];  // MostExpensive: first max wins
        any = true;
      }
      if (any) return worst;
    }
    return 1.7976931348623157e308;
  }
};

static void export_requirements(const Prob& P, const Requirements& reqs, int mask_words, const std::vector<int>& woff,
                                uint8_t* flags, int64_t* gte, int64_t* lte, uint64_t* mask) {
  int K = P.p->n_keys;
  for (int k = 0; k < K; k++) {
    flags[k] = 0;
    gte[k] = 0;
    lte[k] = 0;
  }
  for (int w = 0; w < mask_words; w++) mask[w] = 0;
  for (auto& kv : reqs.m) {
    int k = kv.first;
    if (k == P.hostname_key) continue;
    const Requirement& r = kv.second;
    uint8_t f = KP_SLOT_PRESENT;
    if (r.complement) f |= KP_REQ_COMPLEMENT;
    if (r.has_gte) {
      f |= KP_REQ_HAS_GTE;
      gte[k] = r.gte;
    }
    if (r.has_lte) {
      f |= KP_REQ_HAS_LTE;
      lte[k] = r.lte;
    }
    flags[k] = f;
    for (int32_t v : r.values)
      if (v >= 0 && v < P.nvalues(k)) mask[woff[k] + (v >> 6)] |= 1ull << (v & 63);
  }
}

// reserved offerings without their reservation id / capacity cannot be run through the ReservationManager
static bool reserved_offerings_malformed(const kp_problem* p) {
  if (!p->off_reserved) return false;
  for (int t = 0; t < p->n_its; t++)
    for (int o = p->it_off_off[t]; o < p->it_off_off[t + 1]; o++)
      if (p->off_reserved[o] && (!p->off_reservation_id || !p->off_reservation_capacity || p->off_reservation_id[o] < 0 ||
                                 p->off_reservation_id[o] >= p->n_reservations || p->n_reservations > 64))
        return true;
  return false;
}

static bool has_min_values(const kp_problem* p) {
  for (int e = 0; e < p->n_reqs; e++)
    if (p->req_flags[e] & KP_REQ_HAS_MINVALUES) return true;
  return false;
}

static void fill_result(const Prob& P, Scheduler& s, const std::vector<int32_t>& targets,
                        const std::vector<uint8_t>& errors, kp_result* out) {
  const kp_problem* p = P.p;
  memset(out, 0, sizeof(*out));
  int64_t n = (int64_t)targets.size();
  out->n_pods = n;
  out->pod_target = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
  out->pod_error = (uint8_t*)malloc(n ? n : 1);
  memcpy(out->pod_target, targets.data(), sizeof(int32_t) * n);
  memcpy(out->pod_error, errors.data(), n);
  int C = (int)s.claim_store.size();
  int K = p->n_keys;
  std::vector<int> woff(K + 1, 0);
  for (int k = 0; k < K; k++) woff[k + 1] = woff[k] + (k == P.hostname_key ? 0 : (P.nvalues(k) + 63) / 64);
  int MW = woff[K];
  int ITW = (p->n_its + 63) / 64;
  out->n_claims = C;
  out->n_keys = K;
  out->mask_words = MW;
  out->it_words = ITW;
  size_t c1 = C ? C : 1;
  out->claim_template = (int32_t*)calloc(c1, sizeof(int32_t));
  out->claim_npods = (int32_t*)calloc(c1, sizeof(int32_t));
  out->claim_rank = (int32_t*)calloc(c1, sizeof(int32_t));
  out->claim_requests = (int64_t*)calloc(c1 * P.R, sizeof(int64_t));
  out->claim_its = (uint64_t*)calloc(c1 * (ITW ? ITW : 1), sizeof(uint64_t));
  out->claim_req_flags = (uint8_t*)calloc(c1 * (K ? K : 1), 1);
  out->claim_req_gte = (int64_t*)calloc(c1 * (K ? K : 1), sizeof(int64_t));
  out->claim_req_lte = (int64_t*)calloc(c1 * (K ? K : 1), sizeof(int64_t));
  out->claim_req_mask = (uint64_t*)calloc(c1 * (MW ? MW : 1), sizeof(uint64_t));
  out->claim_reservations = (uint64_t*)calloc(c1, sizeof(uint64_t));
  out->claim_dropped = (uint8_t*)calloc(c1, 1);
  for (int k = 0; k < C; k++)
    for (int id : s.claim_store[k]->reserved) out->claim_reservations[k] |= 1ull << id;
  if (p->max_instance_types > 0) {  // Results.TruncateInstanceTypes (scheduler.go:361-379)
    Pricing pr(P);
    for (int k = 0; k < C; k++) {
      InflightClaim& c = *s.claim_store[k];
      if ((int)c.its.size() <= p->max_instance_types) continue;  // (the order itself is not part of the result)
      pr.order_by_price(c.its, c.reqs);
      c.its.resize(p->max_instance_types);
      if (!p->min_values_best_effort && c.reqs.has_min_values() && !s.satisfies_min_values(c.its, c.reqs)) out->claim_dropped[k] = 1;
    }
    for (int64_t i = 0; i < n; i++)
      if (out->pod_target[i] <= -2 && out->claim_dropped[-2 - out->pod_target[i]]) out->pod_error[i] = KP_PODERR_MINVALUES_TRUNCATED;
  }
  for (size_t pos = 0; pos < s.new_claims.size(); pos++) out->claim_rank[s.new_claims[pos]->created] = (int32_t)pos;
  for (int k = 0; k < C; k++) {
    InflightClaim& c = *s.claim_store[k];
    out->claim_template[k] = c.tmpl;
    out->claim_npods[k] = (int32_t)c.pods.size();
    for (int r = 0; r < P.R; r++) out->claim_requests[(size_t)k * P.R + r] = c.requests.v[r];
    for (int it : c.its) out->claim_its[(size_t)k * ITW + (it >> 6)] |= 1ull << (it & 63);
    export_requirements(P, c.reqs, MW, woff, out->claim_req_flags + (size_t)k * K, out->claim_req_gte + (size_t)k * K,
                        out->claim_req_lte + (size_t)k * K, out->claim_req_mask + (size_t)k * MW);
  }
  // topology counters of the non-hostname groups, group order = creation order (regular groups then inverse)
  std::vector<int32_t> off{0}, cnt;
  auto dump = [&](TopologyGroup& g) {
    if (g.key != P.hostname_key) {
      int nv = P.nvalues(g.key);
      size_t base = cnt.size();
      cnt.resize(base + nv, 0);
      for (auto& kv : g.domains)
        if (kv.first < nv) cnt[base + kv.first] = kv.second;
    }
    off.push_back((int32_t)cnt.size());
  };
  for (auto& g : s.topo.groups) dump(*g);
  for (auto& g : s.topo.inverse_groups) dump(*g);
  out->n_groups = (int32_t)off.size() - 1;
  out->n_domain_slots = (int32_t)cnt.size();
  out->group_domain_off = (int32_t*)malloc(sizeof(int32_t) * off.size());
  memcpy(out->group_domain_off, off.data(), sizeof(int32_t) * off.size());
  out->domain_counts = (int32_t*)malloc(sizeof(int32_t) * (cnt.size() ? cnt.size() : 1));
  if (!cnt.empty()) memcpy(out->domain_counts, cnt.data(), sizeof(int32_t) * cnt.size());
  out->n_existing_evals = s.ctr.existing;
  out->n_inflight_evals = s.ctr.inflight;
  out->n_template_evals = s.ctr.tmpl;
  out->n_commits = s.ctr.commits;
}

}  // namespace orc

using namespace orc;

extern "C" {

int orc_version(void) { return KP_ABI_VERSION; }

int orc_solve_mt(const kp_problem* p, kp_result* out, int threads);
int orc_solve(const kp_problem* p, kp_result* out) { return orc_solve_mt(p, out, 1); }

// threads > 1: in-flight candidates are evaluated by a worker pool like the reference's parallelizeUntil
int orc_solve_mt(const kp_problem* p, kp_result* out, int threads) {
  if (reserved_offerings_malformed(p)) return KP_ERR_INVALID;
  if (p->n_resources > KP_MAX_RESOURCES) return KP_ERR_CAPACITY;
  Prob P(p);
  Scheduler s(P);
  std::unique_ptr<WorkerPool> pool;
  if (threads > 1) {
    pool.reset(new WorkerPool(threads));
    s.pool = pool.get();
  }
  std::vector<uint8_t> active(p->n_nodes, 0);
  for (int i = 0; i < p->n_nodes; i++) active[i] = (p->node_flags[i] & KP_NODE_SCHEDULABLE) != 0;
  std::vector<std::pair<int, int>> bound;
  for (int64_t i = 0; i < p->n_running; i++) bound.push_back({p->run_class[i], p->run_node[i]});
  std::vector<int> pending;
  std::vector<int64_t> rows(p->n_pods);
  for (int64_t i = 0; i < p->n_pods; i++) {
    rows[i] = i;
    pending.push_back(p->pod_class[i]);
  }
  auto t0 = std::chrono::steady_clock::now();
  s.init(active, bound, pending);
  std::vector<int32_t> targets;
  std::vector<uint8_t> errors;
  s.solve(rows, &targets, &errors);
  auto t1 = std::chrono::steady_clock::now();
  fill_result(P, s, targets, errors, out);
  out->solve_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  return KP_OK;
}

void orc_result_free(kp_result* r) {
  free(r->pod_target);
  free(r->pod_error);
  free(r->claim_template);
  free(r->claim_npods);
  free(r->claim_rank);
  free(r->claim_requests);
  free(r->claim_its);
  free(r->claim_req_flags);
  free(r->claim_req_gte);
  free(r->claim_req_lte);
  free(r->claim_req_mask);
  free(r->group_domain_off);
  free(r->domain_counts);
  free(r->claim_reservations);
  free(r->claim_dropped);
  memset(r, 0, sizeof(*r));
}

// filterInstanceTypesByRequirements for a fresh NodeClaim of (class, template) without topology -- parity target of K1
int orc_feasibility(const kp_problem* p, uint64_t* out_bits, int32_t* out_it_words) {
  Prob P(p);
  Scheduler s(P);
  int ITW = (p->n_its + 63) / 64;
  *out_it_words = ITW;
  Res zero;
  for (int cls = 0; cls < p->n_classes; cls++)
    for (int n = 0; n < p->n_templates; n++) {
      uint64_t* row = out_bits + ((size_t)cls * p->n_templates + n) * ITW;
      for (int w = 0; w < ITW; w++) row[w] = 0;
      std::vector<int> its(p->tmpl_its + p->tmpl_it_off[n], p->tmpl_its + p->tmpl_it_off[n + 1]);
      its = s.filter_instance_types(its, P.reqsets[p->tmpl_reqset[n]], zero);
      if (!P.tolerates(p->tmpl_taintset[n], p->class_tolset[cls])) continue;
      Requirements reqs = P.reqsets[p->tmpl_reqset[n]];
      const Requirements& pod_reqs = P.reqsets[p->class_reqset[cls]];
      if (!compatible(P, P, reqs, pod_reqs, true)) continue;
      reqs.add_all(P, pod_reqs);
      Res total = merge(P.R, P.res_row(p->tmpl_daemon, n, P.all_mask()), s.class_requests(cls));
      for (int it : s.filter_instance_types(its, reqs, total)) row[it >> 6] |= 1ull << (it & 63);
    }
  return KP_OK;
}

// disruption/helpers.go:51-142 + consolidation.go:136-229, one call per subset
int orc_consolidate_mt(const kp_problem* p, const kp_consol_input* in, kp_consol_result* out, int threads);
int orc_consolidate(const kp_problem* p, const kp_consol_input* in, kp_consol_result* out) {
  return orc_consolidate_mt(p, in, out, 1);
}
// subsets are independent simulations: `threads` workers take them round-robin
int orc_consolidate_mt(const kp_problem* p, const kp_consol_input* in, kp_consol_result* out, int threads) {
  // BestEffort lowers minValues per NodeClaim during the simulation (nodeclaim.go:186-191); carrying those per-claim values
  // through the decision is not built: refused, like the CUDA path does
  if (has_min_values(p) && p->min_values_best_effort) return KP_ERR_UNSUPPORTED;
  if (reserved_offerings_malformed(p)) return KP_ERR_INVALID;
  Prob P(p);
  Pricing pr(P);
  int ITW = (p->n_its + 63) / 64;
  memset(out, 0, sizeof(*out));
  int S = in->n_subsets;
  out->n_subsets = S;
  out->it_words = ITW;
  out->decision = (uint8_t*)calloc(S ? S : 1, 1);
  out->replacement_its = (uint64_t*)calloc((size_t)(S ? S : 1) * (ITW ? ITW : 1), sizeof(uint64_t));
  out->n_new_claims = (int32_t*)calloc(S ? S : 1, sizeof(int32_t));
  out->n_unscheduled = (int32_t*)calloc(S ? S : 1, sizeof(int32_t));
  // the replacement NodeClaim of a REPLACE (consolidation.go:206-229)
  const int K = p->n_keys;
  std::vector<int> woff(K + 1, 0);
  for (int k = 0; k < K; k++) woff[k + 1] = woff[k] + (k == P.hostname_key ? 0 : (P.nvalues(k) + 63) / 64);
  const int MW = woff[K];
  const size_t s1 = S ? S : 1;
  out->n_keys = K;
  out->mask_words = MW;
  out->n_resources = P.R;
  out->repl_template = (int32_t*)malloc(s1 * 4);
  for (size_t i = 0; i < s1; i++) out->repl_template[i] = -1;
  out->repl_requests = (int64_t*)calloc(s1 * (P.R ? P.R : 1), 8);
  out->repl_req_flags = (uint8_t*)calloc(s1 * (K ? K : 1), 1);
  out->repl_req_gte = (int64_t*)calloc(s1 * (K ? K : 1), 8);
  out->repl_req_lte = (int64_t*)calloc(s1 * (K ? K : 1), 8);
  out->repl_req_mask = (uint64_t*)calloc(s1 * (MW ? MW : 1), 8);
  std::vector<std::vector<int32_t>> order_rows(in->export_price_order ? S : 0);
  const int n_extra = in->n_extra_pods > 0 ? in->n_extra_pods : 0;
  const int64_t extra_row0 = p->n_nodes > 0 ? in->node_pod_off[p->n_nodes] : 0;
  if (n_extra > 0 && (!in->extra_pod_kind || extra_row0 + n_extra != p->n_pods)) return KP_ERR_INVALID;
  const int ct_order[3] = {in->ct_reserved, in->ct_spot, in->ct_on_demand};
  auto t0 = std::chrono::steady_clock::now();
  auto one_subset = [&](int s_i) {
    std::vector<uint8_t> active(p->n_nodes, 0), is_cand(p->n_nodes, 0);
    for (int i = in->subset_off[s_i]; i < in->subset_off[s_i + 1]; i++) is_cand[in->subset_nodes[i]] = 1;
    for (int i = 0; i < p->n_nodes; i++) active[i] = (p->node_flags[i] & KP_NODE_SCHEDULABLE) && !is_cand[i];
    std::vector<std::pair<int, int>> bound;
    for (int64_t i = 0; i < p->n_running; i++) bound.push_back({p->run_class[i], p->run_node[i]});
    std::vector<int64_t> rows;
    std::vector<int> pending;
    for (int n = 0; n < p->n_nodes; n++)
      for (int j = in->node_pod_off[n]; j < in->node_pod_off[n + 1]; j++) {
        if (is_cand[n]) {
          rows.push_back(j);
          pending.push_back(p->pod_class[j]);
        } else {
          bound.push_back({p->pod_class[j], n});  // still running where it is
        }
      }
    const size_t n_cand_pods = rows.size();
    for (int i = 0; i < n_extra; i++) {  // pending pods + reschedulable pods of deleting nodes (helpers.go:65-91)
      rows.push_back(extra_row0 + i);
      pending.push_back(p->pod_class[extra_row0 + i]);
    }
    Scheduler sch(P);
    sch.init(active, bound, pending);
    std::vector<int32_t> targets;
    std::vector<uint8_t> errors;
    sch.solve(rows, &targets, &errors);
    int unscheduled = 0;
    for (size_t i = 0; i < targets.size(); i++) {
      const int kind = i < n_cand_pods ? 0 : in->extra_pod_kind[i - n_cand_pods];
      if (targets[i] == KP_TARGET_UNSCHEDULED) {
        if (kind != KP_EXTRA_PENDING) unscheduled++;  // AllNonPendingPodsScheduled (scheduler.go:330-334)
      } else if (targets[i] >= 0 && !(p->node_flags[targets[i]] & KP_NODE_INITIALIZED)) {
        if (kind == 0) unscheduled++;  // helpers.go:121-140 UninitializedNodeError, not for pods of deleting nodes
      }
    }
    int n_new = (int)sch.claim_store.size();
    out->n_new_claims[s_i] = n_new;
    out->n_unscheduled[s_i] = unscheduled;
    uint8_t decision = KP_DECISION_NOOP;
    uint64_t* rep = out->replacement_its + (size_t)s_i * ITW;
    do {
      if (unscheduled) break;              // consolidation.go:149-155
      if (n_new == 0) {                    // :158-163
        decision = KP_DECISION_DELETE;
        break;
      }
      if (n_new != 1) break;               // :166-171
      InflightClaim& c = *sch.claim_store[0];
      // SimulateScheduling: TruncateInstanceTypes(MaxInstanceTypes = 600) (helpers.go:120, scheduler.go:361-379)
      pr.order_by_price(c.its, c.reqs);
      if (c.its.size() > 600) {
        c.its.resize(600);
        // Truncate (types.go:339-351): the 600 cheapest must still satisfy minValues, else TruncateInstanceTypes drops the
        // NodeClaim and its pods become PodErrors (scheduler.go:361-379) -- not all pods scheduled, nothing to do
        if (c.reqs.has_min_values() && !sch.satisfies_min_values(c.its, c.reqs)) {
          out->n_unscheduled[s_i] = (int32_t)c.pods.size();
          break;
        }
      }
      // getCandidatePrices (consolidation.go:319-337)
      double price = 0;
      bool zero = false;
      bool all_spot = true;
      for (int i = in->subset_off[s_i]; i < in->subset_off[s_i + 1]; i++) {
        int node = in->subset_nodes[i];
        if (!in->node_is_spot[node]) all_spot = false;
        int it = in->node_it[node];
        if (it < 0) {
          zero = true;
          break;
        }
        const Requirements& labels = P.reqsets[p->node_reqset[node]];
        bool any = false;
        double cheapest = 0;
        for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++) {
          if (!compatible(P, P, labels, P.reqsets[p->off_reqset[o]], true)) continue;
          if (!any || p->off_price[o] < cheapest) cheapest = p->off_price[o];
          any = true;
        }
        if (!any) {
          zero = true;
          break;
        }
        price += cheapest;
      }
      if (zero) price = 0.0;
      pr.order_by_price(c.its, c.reqs);  // consolidation.go:186
      Requirement ctreq = in->capacity_type_key >= 0 ? c.reqs.get(in->capacity_type_key) : exists_requirement(0);
      bool spot_ok = in->capacity_type_key >= 0 && in->ct_spot >= 0 && has(P, ctreq, in->ct_spot);
      if (all_spot && spot_ok) {  // computeSpotToSpotConsolidation (consolidation.go:236-316)
        if (!in->spot_to_spot_enabled) break;
        Requirement r;
        r.key = in->capacity_type_key;
        r.values.push_back(in->ct_spot);
        c.reqs.add(P, r);
        std::vector<int> f;
        for (int it : c.its) {  // InstanceTypes.Compatible (types.go:259-267)
          bool ok = false;
          for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1] && !ok; o++)
            ok = p->off_available[o] && compatible(P, P, c.reqs, P.reqsets[p->off_reqset[o]], true);
          if (ok) f.push_back(it);
        }
        c.its.swap(f);
      }
      std::vector<int> kept;  // RemoveInstanceTypeOptionsByPriceAndMinValues (nodeclaim.go:309-318)
      for (int it : c.its)
        if (pr.worst_launch_price(it, c.reqs, in->capacity_type_key, ct_order) < price) kept.push_back(it);
      // ... and SatisfiesMinValues of what is left (nodeclaim.go:314-316): an error is "Filtering by price", no command
      if (c.reqs.has_min_values() && !sch.satisfies_min_values(kept, c.reqs)) break;
      if (kept.empty()) break;
      if (all_spot && spot_ok && (in->subset_off[s_i + 1] - in->subset_off[s_i]) == 1) {
        if (kept.size() < 15) break;  // MinInstanceTypesForSpotToSpotConsolidation
        // the 15 cheapest go out -- or as many as minValues needs, if that is more (consolidation.go:296-312; the shortest
        // prefix of the price order that satisfies every key, types.go:301-337)
        size_t cap = 15;
        if (c.reqs.has_min_values()) {
          std::vector<int> prefix;
          for (size_t i = 0; i < kept.size(); i++) {
            prefix.push_back(kept[i]);
            if (sch.satisfies_min_values(prefix, c.reqs)) {
              cap = std::max<size_t>(15, i + 1);
              break;
            }
          }
        }
        kept.resize(std::min(cap, kept.size()));
      }
      if (in->filter_same_instance_type && in->subset_off[s_i + 1] - in->subset_off[s_i] >= 2) {
        // filterOutSameInstanceType (multinodeconsolidation.go:189-226)
        std::map<int, double> price_by_type;  // cheapest node of each instance type that is being removed
        std::set<int> existing;
        for (int i = in->subset_off[s_i]; i < in->subset_off[s_i + 1]; i++) {
          int node = in->subset_nodes[i], it = in->node_it[node];
          if (it < 0) continue;
          existing.insert(it);
          const Requirements& labels = P.reqsets[p->node_reqset[node]];
          bool any = false;
          double cheapest = 0;
          for (int o = p->it_off_off[it]; o < p->it_off_off[it + 1]; o++) {
            if (!compatible(P, P, labels, P.reqsets[p->off_reqset[o]], true)) continue;
            if (!any || p->off_price[o] < cheapest) cheapest = p->off_price[o];
            any = true;
          }
          if (!any) continue;
          auto f = price_by_type.find(it);
          if (f == price_by_type.end() || cheapest < f->second) price_by_type[it] = cheapest;
        }
        double max_price = std::numeric_limits<double>::max();
        for (int it : kept)
          if (existing.count(it)) {
            auto f = price_by_type.find(it);
            double v = f == price_by_type.end() ? 0.0 : f->second;  // a Go map miss reads as 0
            if (v < max_price) max_price = v;
          }
        std::vector<int> kept2;
        for (int it : kept)
          if (pr.worst_launch_price(it, c.reqs, in->capacity_type_key, ct_order) < max_price) kept2.push_back(it);
        kept.swap(kept2);
        // RemoveInstanceTypeOptionsByPriceAndMinValues again (multinodeconsolidation.go:220-224): an error or an empty list
        // is not a valid command for the binary search (:157-163)
        if (c.reqs.has_min_values() && !sch.satisfies_min_values(kept, c.reqs)) break;
        if (kept.empty()) break;
      }
      decision = KP_DECISION_REPLACE;
      for (int it : kept) rep[it >> 6] |= 1ull << (it & 63);
      // OD -> [OD, spot]: pin the replacement to spot (consolidation.go:206-214); the spot-to-spot path pinned it above
      if (!(all_spot && spot_ok) && in->capacity_type_key >= 0 && in->ct_spot >= 0 && in->ct_on_demand >= 0) {
        Requirement cur = c.reqs.get(in->capacity_type_key);
        if (has(P, cur, in->ct_spot) && has(P, cur, in->ct_on_demand)) {
          Requirement r;
          r.key = in->capacity_type_key;
          r.values.push_back(in->ct_spot);
          c.reqs.add(P, r);
        }
      }
      out->repl_template[s_i] = c.tmpl;
      for (int r = 0; r < P.R; r++) out->repl_requests[(size_t)s_i * P.R + r] = c.requests.v[r];
      export_requirements(P, c.reqs, MW, woff, out->repl_req_flags + (size_t)s_i * K, out->repl_req_gte + (size_t)s_i * K,
                          out->repl_req_lte + (size_t)s_i * K, out->repl_req_mask + (size_t)s_i * MW);
      if (in->export_price_order) order_rows[s_i].assign(kept.begin(), kept.end());
    } while (0);
    out->decision[s_i] = decision;
  };
  if (threads <= 1) {
    for (int s_i = 0; s_i < S; s_i++) one_subset(s_i);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++)
      pool.emplace_back([&, w] {
        for (int s_i = w; s_i < S; s_i += threads) one_subset(s_i);
      });
    for (auto& t : pool) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  out->solve_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  if (in->export_price_order) {
    out->repl_order_off = (int32_t*)calloc((size_t)S + 1, 4);
    size_t tot = 0;
    for (int s_ = 0; s_ < S; s_++) {
      tot += order_rows[s_].size();
      out->repl_order_off[s_ + 1] = (int32_t)tot;
    }
    out->repl_order = (int32_t*)calloc(tot ? tot : 1, 4);
    for (int s_ = 0; s_ < S; s_++)
      if (!order_rows[s_].empty()) memcpy(out->repl_order + out->repl_order_off[s_], order_rows[s_].data(), order_rows[s_].size() * 4);
  }
  return KP_OK;
}

void orc_consol_result_free(kp_consol_result* r) {
  free(r->decision);
  free(r->replacement_its);
  free(r->n_new_claims);
  free(r->n_unscheduled);
  free(r->repl_template);
  free(r->repl_requests);
  free(r->repl_req_flags);
  free(r->repl_req_gte);
  free(r->repl_req_lte);
  free(r->repl_req_mask);
  free(r->repl_order_off);
  free(r->repl_order);
  memset(r, 0, sizeof(*r));
}

// ---- known-answer-test hooks for the requirement algebra (tests/test_oracle_requirements.py) ----
// A requirement is passed as: flags (KP_REQ_*), gte, lte, min_values, n values. The int table for the single key is
// (value_int, value_is_int) of n_universe values.
struct KatTable : IntTable {
  const int64_t* vi;
  const uint8_t* is;
  int n;
  bool atoi(int, int32_t v, int64_t* out) const override {
    if (v < 0 || v >= n || !is[v]) return false;
    *out = vi[v];
    return true;
  }
};
static Requirement kat_req(uint8_t flags, int64_t gte, int64_t lte, int32_t mv, const int32_t* vals, int n) {
  Requirement r;
  r.key = 0;
  r.complement = flags & KP_REQ_COMPLEMENT;
  r.has_gte = flags & KP_REQ_HAS_GTE;
  r.has_lte = flags & KP_REQ_HAS_LTE;
  r.has_min = flags & KP_REQ_HAS_MINVALUES;
  r.gte = gte;
  r.lte = lte;
  r.min_values = mv;
  for (int i = 0; i < n; i++) r.insert(vals[i]);
  return r;
}
// NewRequirementWithFlexibility: returns canonical form
int orc_kat_new_requirement(int op, int64_t operand, int has_min, int32_t min_values, const int32_t* vals, int n,
                            uint8_t* out_flags, int64_t* out_gte, int64_t* out_lte, int32_t* out_min, int32_t* out_vals,
                            int32_t* out_n) {
  std::vector<int32_t> v(vals, vals + n);
  Requirement r = new_requirement(0, (Operator)op, v, operand, has_min, min_values);
  *out_flags = (r.complement ? KP_REQ_COMPLEMENT : 0) | (r.has_gte ? KP_REQ_HAS_GTE : 0) |
               (r.has_lte ? KP_REQ_HAS_LTE : 0) | (r.has_min ? KP_REQ_HAS_MINVALUES : 0);
  *out_gte = r.gte;
  *out_lte = r.lte;
  *out_min = r.min_values;
  *out_n = (int32_t)r.values.size();
  for (size_t i = 0; i < r.values.size(); i++) out_vals[i] = r.values[i];
  return (int)r.op();
}
int orc_kat_intersection(const int64_t* vi, const uint8_t* is, int n_universe, uint8_t fa, int64_t ga, int64_t la,
                         int32_t ma, const int32_t* va, int na, uint8_t fb, int64_t gb, int64_t lb, int32_t mb,
                         const int32_t* vb, int nb, uint8_t* out_flags, int64_t* out_gte, int64_t* out_lte,
                         int32_t* out_min, int32_t* out_vals, int32_t* out_n, int32_t* out_has_intersection) {
  KatTable t;
  t.vi = vi;
  t.is = is;
  t.n = n_universe;
  Requirement a = kat_req(fa, ga, la, ma, va, na), b = kat_req(fb, gb, lb, mb, vb, nb);
  Requirement r = intersection(t, a, b);
  *out_has_intersection = has_intersection(t, a, b);
  *out_flags = (r.complement ? KP_REQ_COMPLEMENT : 0) | (r.has_gte ? KP_REQ_HAS_GTE : 0) |
               (r.has_lte ? KP_REQ_HAS_LTE : 0) | (r.has_min ? KP_REQ_HAS_MINVALUES : 0);
  *out_gte = r.gte;
  *out_lte = r.lte;
  *out_min = r.min_values;
  *out_n = (int32_t)r.values.size();
  for (size_t i = 0; i < r.values.size(); i++) out_vals[i] = r.values[i];
  return (int)r.op();
}
int orc_kat_has(const int64_t* vi, const uint8_t* is, int n_universe, uint8_t fa, int64_t ga, int64_t la,
                const int32_t* va, int na, int32_t value) {
  KatTable t;
  t.vi = vi;
  t.is = is;
  t.n = n_universe;
  return has(t, kat_req(fa, ga, la, 0, va, na), value);
}
// Requirements.Compatible(A, B, options) on a problem's reqsets (strict when allow_undefined == 0)
int orc_kat_compatible(const kp_problem* p, int reqset_a, int reqset_b, int allow_undefined) {
  Prob P(p);
  return compatible(P, P, P.reqsets[reqset_a], P.reqsets[reqset_b], allow_undefined != 0);
}
// sort.Slice on int keys: writes the permutation Go would leave
void orc_kat_gosort(const int64_t* keys, int n, int32_t* perm) {
  for (int i = 0; i < n; i++) perm[i] = i;
  go_sort_slice(n, [&](int a, int b) { return keys[perm[a]] < keys[perm[b]]; },
                [&](int a, int b) { std::swap(perm[a], perm[b]); });
}
}
