// TEST INFRASTRUCTURE -- CPU oracle (see orc_requirement.hpp header).
//
// orc_model.hpp: decoding of the flat kp_problem (include/karpsolve.h) into the reference's object model:
// scheduling.Requirements, corev1.ResourceList, taints / tolerations, label selectors.
#pragma once
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../include/karpsolve.h"
#include "orc_requirement.hpp"

namespace orc {

// corev1.ResourceList as a dense vector with a presence mask (map-key semantics matter for Subtract / limits)
struct Res {
  int64_t v[KP_MAX_RESOURCES];
  uint32_t present;
  Res() : present(0) { memset(v, 0, sizeof(v)); }
};

struct Prob : IntTable, KeyPolicy {
  const kp_problem* p;
  int R;
  int hostname_key = -1;
  int cpu_res = -1, mem_res = -1, nodes_res = -1;
  std::vector<Requirements> reqsets;  // decoded + folded
  std::vector<Res> it_alloc;          // InstanceType.Allocatable() (types.go:198-236)

  explicit Prob(const kp_problem* pp) : p(pp), R(pp->n_resources) {
    for (int k = 0; k < p->n_keys; k++)
      if (p->key_flags[k] & KP_KEY_HOSTNAME) hostname_key = k;
    for (int r = 0; r < R; r++) {
      if (p->res_flags[r] & KP_RES_CPU) cpu_res = r;
      if (p->res_flags[r] & KP_RES_MEMORY) mem_res = r;
      if (p->res_flags[r] & KP_RES_NODES) nodes_res = r;
    }
    reqsets.resize(p->n_reqsets);
    for (int s = 0; s < p->n_reqsets; s++) {
      for (int e = p->reqset_off[s]; e < p->reqset_off[s + 1]; e++) {
        Requirement r;
        r.key = p->req_key[e];
        uint8_t f = p->req_flags[e];
        r.complement = f & KP_REQ_COMPLEMENT;
        r.has_gte = f & KP_REQ_HAS_GTE;
        r.has_lte = f & KP_REQ_HAS_LTE;
        r.has_min = f & KP_REQ_HAS_MINVALUES;
        r.gte = p->req_gte ? p->req_gte[e] : 0;
        r.lte = p->req_lte ? p->req_lte[e] : 0;
        r.min_values = p->req_min_values ? p->req_min_values[e] : 0;
        for (int i = p->req_val_off[e]; i < p->req_val_off[e + 1]; i++) r.insert(p->req_vals[i]);
        reqsets[s].add(*this, r);
      }
    }
    it_alloc.resize(p->n_its);
    for (int t = 0; t < p->n_its; t++) it_alloc[t] = allocatable(t);
  }
  int nvalues(int key) const { return p->key_value_off[key + 1] - p->key_value_off[key]; }
  bool atoi(int key, int32_t value, int64_t* out) const override {
    if (value < 0 || value >= nvalues(key)) return false;  // NodeClaim hostname placeholders are not integers
    int idx = p->key_value_off[key] + value;
    if (!p->value_is_int[idx]) return false;
    *out = p->value_int[idx];
    return true;
  }
  bool well_known(int key) const override { return p->key_flags[key] & KP_KEY_WELL_KNOWN; }

  Res res_row(const int64_t* base, int idx, uint32_t present) const {
    Res r;
    for (int i = 0; i < R; i++) r.v[i] = base ? base[(int64_t)idx * R + i] : 0;
    r.present = present;
    return r;
  }
  uint32_t all_mask() const { return (1u << R) - 1; }

  // types.go:198-216 precompute(): Subtract(Capacity, Overhead.Total()) keeps Capacity's keys; hugepages come off
  // memory, floored at zero.
  Res allocatable(int t) const {
    Res a;
    uint32_t cp = p->it_cap_present ? p->it_cap_present[t] : all_mask();
    a.present = cp;
    for (int r = 0; r < R; r++) {
      if (!(cp >> r & 1)) continue;
      a.v[r] = p->it_capacity[(int64_t)t * R + r] - (p->it_overhead ? p->it_overhead[(int64_t)t * R + r] : 0);
    }
    for (int r = 0; r < R; r++) {
      if ((cp >> r & 1) && (p->res_flags[r] & KP_RES_HUGEPAGES) && mem_res >= 0) {
        int64_t cur = (a.present >> mem_res & 1) ? a.v[mem_res] : 0;
        cur -= p->it_capacity[(int64_t)t * R + r];
        if (cur < 0) cur = 0;
        a.v[mem_res] = cur;
        a.present |= 1u << mem_res;
      }
    }
    return a;
  }

  // ---- taints (pkg/scheduling/taints.go:49-66 over corev1.Toleration.ToleratesTaint, k8s.io/api v0.35.0) ----
  bool tolerates_taint(int tol, int taint) const {
    uint8_t te = p->tol_effect[tol];
    if (te != KP_EFFECT_NONE && te != p->taint_effect[taint]) return false;
    int tk = p->tol_key[tol];
    if (tk != 0 && tk != p->taint_key[taint]) return false;
    switch (p->tol_op[tol]) {
      case KP_TOL_EQUAL:
        return p->tol_value[tol] == p->taint_value[taint];
      case KP_TOL_EXISTS:
        return true;
      case KP_TOL_LT:
      case KP_TOL_GT: {
        int a = p->taint_value[taint], b = p->tol_value[tol];
        if (!p->tt_is_int || !p->tt_is_int[a] || !p->tt_is_int[b]) return false;
        return p->tol_op[tol] == KP_TOL_LT ? p->tt_int[a] < p->tt_int[b] : p->tt_int[a] > p->tt_int[b];
      }
    }
    return false;
  }
  // Taints.Tolerates(tolerations) == nil
  bool tolerates(int taintset, int tolset) const {
    if (taintset < 0) return true;
    for (int i = p->taintset_off[taintset]; i < p->taintset_off[taintset + 1]; i++) {
      int taint = p->taintset_ids[i];
      bool ok = false;
      if (tolset >= 0)
        for (int j = p->tolset_off[tolset]; j < p->tolset_off[tolset + 1] && !ok; j++)
          ok = tolerates_taint(p->tolset_ids[j], taint);
      if (!ok) return false;
    }
    return true;
  }
  int taintset_size(int ts) const { return ts < 0 ? 0 : p->taintset_off[ts + 1] - p->taintset_off[ts]; }

  // ---- label selectors (metav1.LabelSelectorAsSelector + labels.Selector.Matches, k8s.io/apimachinery) ----
  bool label_lookup(int labelset, int key, int* val) const {
    if (labelset < 0) return false;
    for (int i = p->labelset_off[labelset]; i < p->labelset_off[labelset + 1]; i++)
      if (p->label_key[i] == key) {
        *val = p->label_val[i];
        return true;
      }
    return false;
  }
  bool selector_matches(int selector, int labelset) const {
    if (selector < 0) return false;  // nil selector => labels.Nothing() (topologygroup.go:101-104)
    for (int e = p->selector_off[selector]; e < p->selector_off[selector + 1]; e++) {
      int val = 0;
      bool has = label_lookup(labelset, p->selx_key[e], &val);
      bool in = false;
      if (has)
        for (int i = p->selx_val_off[e]; i < p->selx_val_off[e + 1]; i++)
          if (p->selx_vals[i] == val) in = true;
      switch (p->selx_op[e]) {
        case KP_SEL_IN:
          if (!(has && in)) return false;
          break;
        case KP_SEL_NOT_IN:
          if (has && in) return false;
          break;
        case KP_SEL_EXISTS:
          if (!has) return false;
          break;
        case KP_SEL_DOES_NOT_EXIST:
          if (has) return false;
          break;
      }
    }
    return true;
  }
  bool nsset_has(int nsset, int ns) const {
    for (int i = p->nsset_off[nsset]; i < p->nsset_off[nsset + 1]; i++)
      if (p->nsset_ids[i] == ns) return true;
    return false;
  }
};

// ---- pkg/utils/resources/resources.go ----
// Fits (resources.go:150-163): any negative entry of `total` => false; then candidate <= total (missing => 0)
inline bool fits(int R, const Res& candidate, const Res& total) {
  for (int r = 0; r < R; r++)
    if ((total.present >> r & 1) && total.v[r] < 0) return false;
  for (int r = 0; r < R; r++) {
    int64_t t = (total.present >> r & 1) ? total.v[r] : 0;
    if (candidate.v[r] > t) return false;
  }
  return true;
}
// Merge (resources.go:52-67)
inline Res merge(int R, const Res& a, const Res& b) {
  Res o;
  for (int r = 0; r < R; r++) o.v[r] = a.v[r] + b.v[r];
  o.present = a.present | b.present;
  return o;
}

}  // namespace orc
