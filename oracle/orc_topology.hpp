// TEST INFRASTRUCTURE -- CPU oracle (see orc_requirement.hpp header).
//
// orc_topology.hpp: restatement of pkg/controllers/provisioning/scheduling/{topology.go, topologygroup.go,
// topologynodefilter.go, topologydomaingroup.go}.  Wherever Go ranges over a map this code iterates in
// ascending interned id -- the canonical order the CUDA path mirrors.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <sstream>

#include "orc_model.hpp"

namespace orc {

enum { PolicyIgnore = 0, PolicyHonor = 1, PolicyUnset = 2 };

// topologynodefilter.go:31-36
struct NodeFilter {
  std::vector<Requirements> reqs;
  int taint_policy = PolicyUnset, affinity_policy = PolicyUnset;
  int tolset = -1;

  // topologynodefilter.go:68-97 (note: compatibility options are NOT forwarded to matchesRequirements, :72 vs :85)
  bool matches(const Prob& P, int taintset, const Requirements& requirements) const {
    bool matches_affinity = true;
    if (affinity_policy == PolicyHonor) {
      if (!reqs.empty()) {
        matches_affinity = false;
        for (auto& r : reqs)
          if (compatible(P, P, requirements, r, false)) {
            matches_affinity = true;
            break;
          }
      }
    }
    bool matches_taints = true;
    if (taint_policy == PolicyHonor) matches_taints = P.tolerates(taintset, tolset);
    return matches_affinity && matches_taints;
  }
};

// topologydomaingroup.go:28-72: domain -> list of taint sets it is reachable under
struct DomainGroup {
  std::map<int32_t, std::vector<int>> d;  // taintset ids; -1 or an empty set == "no taints"
  void insert(const Prob& P, int32_t domain, int taintset) {
    bool empty = P.taintset_size(taintset) == 0;
    auto it = d.find(domain);
    if (it == d.end() || empty) {
      d[domain] = {taintset};
      return;
    }
    if (P.taintset_size(it->second[0]) == 0) return;
    it->second.push_back(taintset);
  }
  template <class F>
  void for_each_domain(const Prob& P, int tolset, int taint_policy, F f) const {
    for (auto& kv : d) {
      if (taint_policy == PolicyIgnore) {
        f(kv.first);
        continue;
      }
      for (int ts : kv.second)
        if (P.tolerates(ts, tolset)) {
          f(kv.first);
          break;
        }
    }
  }
};

struct TopologyGroup {
  int key = 0, type = 0;
  int32_t max_skew = 0;
  bool has_min_domains = false;
  int32_t min_domains = 0;
  int nsset = -1;     // namespaces
  int selector = -1;  // rawSelector
  NodeFilter filter;
  std::set<int> owners;  // pod classes that have this topology as a scheduling rule
  std::map<int32_t, int32_t> domains;
  std::set<int32_t> empty_domains;
  int index = 0;  // creation order
  bool inverse = false;

  // topologygroup.go:431-433
  bool selects(const Prob& P, int cls) const {
    return P.nsset_has(nsset, P.p->class_namespace[cls]) && P.selector_matches(selector, P.p->class_labelset[cls]);
  }
  // topologygroup.go:148-150
  bool counts(const Prob& P, int cls, int taintset, const Requirements& reqs) const {
    return selects(P, cls) && filter.matches(P, taintset, reqs);
  }
  void record(int32_t domain) {
    domains[domain]++;
    empty_domains.erase(domain);
  }
  void register_domain(int32_t domain) {
    if (!domains.count(domain)) {
      domains[domain] = 0;
      empty_domains.insert(domain);
    }
  }
  int32_t count_of(int32_t domain) const {
    auto it = domains.find(domain);
    return it == domains.end() ? 0 : it->second;
  }

  // topologygroup.go:289-310
  int32_t domain_min_count(const Prob& P, const Requirement& pod_domains) const {
    if (key == P.hostname_key) return 0;
    int32_t mn = INT32_MAX;
    int32_t supported = 0;
    for (auto& kv : domains)
      if (has(P, pod_domains, kv.first)) {
        supported++;
        if (kv.second < mn) mn = kv.second;
      }
    if (has_min_domains && supported < min_domains) mn = 0;
    return mn;
  }

  static Requirement in_one(int key, int32_t v) {
    Requirement r;
    r.key = key;
    r.values.push_back(v);
    return r;
  }
  static Requirement does_not_exist(int key) {
    Requirement r;
    r.key = key;
    return r;
  }

  // topologygroup.go:226-287
  Requirement next_domain_spread(const Prob& P, int cls, const Requirement& pod_domains,
                                 const Requirement& node_domains) const {
    int32_t mn = domain_min_count(P, pod_domains);
    bool self = selects(P, cls);
    bool found = false;
    int32_t min_domain = 0;
    int32_t min_count = INT32_MAX;
    if (key == P.hostname_key && node_domains.values.size() == 1) {
      int32_t host = node_domains.values[0];
      int32_t c = count_of(host);
      if (self) c++;
      if (c <= max_skew) return in_one(key, host);
      return does_not_exist(key);
    }
    if (node_domains.op() == OpIn) {
      for (int32_t d : node_domains.values) {  // canonical: ascending id (Go: UnsortedList)
        auto it = domains.find(d);
        if (it == domains.end()) continue;
        int32_t c = it->second;
        if (self) c++;
        if (c - mn <= max_skew && c < min_count) {
          found = true;
          min_domain = d;
          min_count = c;
        }
      }
    } else {
      for (auto& kv : domains) {  // canonical: ascending id (Go: map range)
        if (!has(P, node_domains, kv.first)) continue;
        int32_t c = kv.second;
        if (self) c++;
        if (c - mn <= max_skew && c < min_count) {
          found = true;
          min_domain = kv.first;
          min_count = c;
        }
      }
    }
    if (!found) return does_not_exist(key);
    return in_one(key, min_domain);
  }

  bool any_compatible_pod_domain(const Prob& P, const Requirement& pod_domains) const {
    for (auto& kv : domains)
      if (has(P, pod_domains, kv.first) && kv.second > 0) return true;
    return false;
  }

  // topologygroup.go:313-377
  Requirement next_domain_affinity(const Prob& P, int cls, const Requirement& pod_domains,
                                   const Requirement& node_domains) const {
    Requirement options = does_not_exist(key);
    if (key == P.hostname_key && node_domains.values.size() == 1) {
      int32_t host = node_domains.values[0];
      if (!has(P, pod_domains, host)) return options;
      if (count_of(host) > 0) {
        options.insert(host);
        return options;
      }
      if (selects(P, cls) &&
          (domains.size() == empty_domains.size() || !any_compatible_pod_domain(P, pod_domains))) {
        options.insert(host);
        return options;
      }
      return options;
    }
    if (node_domains.op() == OpIn) {
      for (int32_t d : node_domains.values) {
        auto it = domains.find(d);
        if (has(P, pod_domains, d) && it != domains.end() && it->second > 0) options.insert(d);
      }
    } else {
      for (auto& kv : domains)
        if (has(P, pod_domains, kv.first) && kv.second > 0 && has(P, node_domains, kv.first)) options.insert(kv.first);
    }
    if (options.len() != 0) return options;
    if (selects(P, cls) && (domains.size() == empty_domains.size() || !any_compatible_pod_domain(P, pod_domains))) {
      Requirement intersected = intersection(P, pod_domains, node_domains);
      for (auto& kv : domains)  // canonical "first random domain" == lowest id (topologygroup.go:361-366)
        if (has(P, intersected, kv.first)) {
          options.insert(kv.first);
          break;
        }
      for (auto& kv : domains)  // both loops run (topologygroup.go:369-374)
        if (has(P, pod_domains, kv.first)) {
          options.insert(kv.first);
          break;
        }
    }
    return options;
  }

  // topologygroup.go:393-428
  Requirement next_domain_anti_affinity(const Prob& P, const Requirement& pod_domains,
                                        const Requirement& node_domains) const {
    Requirement options = does_not_exist(key);
    if (key == P.hostname_key && node_domains.values.size() == 1) {
      int32_t host = node_domains.values[0];
      if (count_of(host) == 0) options.insert(host);
      return options;
    }
    if (node_domains.op() == OpIn && node_domains.len() < (int64_t)empty_domains.size()) {
      for (int32_t d : node_domains.values)
        if (empty_domains.count(d) && has(P, pod_domains, d)) options.insert(d);
    } else {
      for (int32_t d : empty_domains)
        if (has(P, node_domains, d) && has(P, pod_domains, d)) options.insert(d);
    }
    return options;
  }

  // topologygroup.go:128-139
  Requirement get(const Prob& P, int cls, const Requirement& pod_domains, const Requirement& node_domains) const {
    switch (type) {
      case KP_TOPO_SPREAD:
        return next_domain_spread(P, cls, pod_domains, node_domains);
      case KP_TOPO_AFFINITY:
        return next_domain_affinity(P, cls, pod_domains, node_domains);
      default:
        return next_domain_anti_affinity(P, pod_domains, node_domains);
    }
  }
};

inline void ser_req(std::ostringstream& o, const Requirements& r) {
  o << "{";
  for (auto& kv : r.m) {
    const Requirement& q = kv.second;
    o << q.key << ":" << q.complement << ":";
    for (auto v : q.values) o << v << ",";
    if (q.has_gte) o << "g" << q.gte;
    if (q.has_lte) o << "l" << q.lte;
    if (q.has_min) o << "m" << q.min_values;
    o << ";";
  }
  o << "}";
}

// A view of the cluster node list the topology needs (labels / taints / hostname)
struct TopoNodeView {
  int n = 0;
  std::vector<uint8_t> active;  // node takes part in this simulation as an existing node
};

struct Topology {
  const Prob& P;
  std::vector<std::unique_ptr<TopologyGroup>> groups;          // topologyGroups, creation order
  std::vector<std::unique_ptr<TopologyGroup>> inverse_groups;  // inverseTopologyGroups
  std::map<std::string, int> group_index, inverse_index;       // Hash() -> index
  std::map<int, DomainGroup> domain_groups;                    // key -> universe
  // which classes have been Update()d (class-level ownership: every pod of a class carries the same constraints)
  std::set<int> updated;

  explicit Topology(const Prob& p) : P(p) {}

  // topology.go:105-143 buildDomainGroups. Deviation (documented in DESIGN.md): the template requirement set
  // (tmpl_reqset) already includes the karpenter.sh/nodepool + nodeclass labels that NewNodeClaimTemplate adds.
  void build_domain_groups() {
    const kp_problem* p = P.p;
    for (int n = 0; n < p->n_templates; n++) {
      const Requirements& np = P.reqsets[p->tmpl_reqset[n]];
      int ts = p->tmpl_taintset[n];
      for (int i = p->tmpl_it_off[n]; i < p->tmpl_it_off[n + 1]; i++) {
        int it = p->tmpl_its[i];
        Requirements r = np;
        r.add_all(P, P.reqsets[p->it_reqset[it]]);
        for (auto& kv : r.m)
          for (int32_t d : kv.second.values) domain_groups[kv.first].insert(P, d, ts);
      }
      for (auto& kv : np.m)
        if (kv.second.op() == OpIn)
          for (int32_t d : kv.second.values) domain_groups[kv.first].insert(P, d, ts);
    }
  }

  // TopologyGroup.Hash() (topologygroup.go:186-220): identity by content
  std::string hash_of(const TopologyGroup& g) const {
    const kp_problem* p = P.p;
    std::ostringstream o;
    o << g.key << "|" << g.type << "|" << g.max_skew << "|ns:";
    std::set<int> ns;
    for (int i = p->nsset_off[g.nsset]; i < p->nsset_off[g.nsset + 1]; i++) ns.insert(p->nsset_ids[i]);
    for (int x : ns) o << x << ",";
    o << "|f:" << g.filter.taint_policy << g.filter.affinity_policy << "[";
    std::set<std::string> rs;
    for (auto& r : g.filter.reqs) {
      std::ostringstream q;
      ser_req(q, r);
      rs.insert(q.str());
    }
    for (auto& s : rs) o << s;
    o << "]t:";
    if (g.filter.tolset >= 0) {
      std::set<std::string> ts;
      for (int j = p->tolset_off[g.filter.tolset]; j < p->tolset_off[g.filter.tolset + 1]; j++) {
        int t = p->tolset_ids[j];
        std::ostringstream q;
        q << p->tol_key[t] << "/" << (int)p->tol_op[t] << "/" << p->tol_value[t] << "/" << (int)p->tol_effect[t];
        ts.insert(q.str());
      }
      for (auto& s : ts) o << s << ",";
    }
    o << "|s:";
    if (g.selector < 0)
      o << "nil";
    else {
      std::set<std::string> ex;
      for (int e = p->selector_off[g.selector]; e < p->selector_off[g.selector + 1]; e++) {
        std::ostringstream q;
        q << p->selx_key[e] << "/" << (int)p->selx_op[e] << "/";
        std::set<int> vs;
        for (int i = p->selx_val_off[e]; i < p->selx_val_off[e + 1]; i++) vs.insert(p->selx_vals[i]);
        for (int v : vs) q << v << ",";
        ex.insert(q.str());
      }
      for (auto& s : ex) o << s << ";";
    }
    return o.str();
  }

  // topologygroup.go:75-126 NewTopologyGroup
  std::unique_ptr<TopologyGroup> new_group(int cls, int c) const {
    const kp_problem* p = P.p;
    auto g = std::make_unique<TopologyGroup>();
    g->type = p->tsc_type[c];
    g->key = p->tsc_key[c];
    g->nsset = p->tsc_nsset[c];
    g->selector = p->tsc_selector[c];
    if (g->type == KP_TOPO_SPREAD) {
      g->max_skew = p->tsc_max_skew[c];
      g->has_min_domains = p->tsc_min_domains[c] >= 0;
      g->min_domains = p->tsc_min_domains[c];
      // MakeTopologyNodeFilter (topologynodefilter.go:38-64); defaults: taints Ignore, affinity Honor
      g->filter.taint_policy = p->tsc_taint_policy[c] ? PolicyHonor : PolicyIgnore;
      g->filter.affinity_policy = p->tsc_affinity_policy[c] ? PolicyHonor : PolicyIgnore;
      g->filter.tolset = p->class_tolset[cls];
      for (int i = p->class_filter_off[cls]; i < p->class_filter_off[cls + 1]; i++)
        g->filter.reqs.push_back(P.reqsets[p->class_filter_reqsets[i]]);
    } else {
      g->max_skew = INT32_MAX;  // math.MaxInt32 (topology.go:311,492)
    }
    auto dg = domain_groups.find(g->key);
    if (dg != domain_groups.end())
      dg->second.for_each_domain(P, p->class_tolset[cls], g->filter.taint_policy, [&](int32_t d) {
        g->domains[d] = 0;
        g->empty_domains.insert(d);
      });
    return g;
  }

  // node label value for a key: labels as In{value}; hostname via node_hostname (topology.go:405-415)
  bool node_domain(int node, int key, int32_t* out) const {
    const kp_problem* p = P.p;
    if (key == P.hostname_key) {
      *out = p->node_hostname[node];
      return true;
    }
    const Requirements& r = P.reqsets[p->node_reqset[node]];
    auto it = r.m.find(key);
    if (it == r.m.end() || it->second.values.empty()) return false;
    *out = it->second.values[0];
    return true;
  }
  Requirements node_label_reqs(int node) const {
    Requirements r = P.reqsets[P.p->node_reqset[node]];
    if (P.hostname_key >= 0) {
      Requirement h;
      h.key = P.hostname_key;
      h.values.push_back(P.p->node_hostname[node]);
      r.add(P, h);
    }
    return r;
  }

  // bound pods visible to the topology: (class, node). Supplied by the scheduler wrapper.
  std::vector<std::pair<int, int>> bound_pods;
  std::vector<uint8_t> state_node;  // node is in stateNodes (t.stateNodes)

  // topology.go:328-426 countDomains
  void count_domains(TopologyGroup& tg) const {
    const kp_problem* p = P.p;
    for (int n = 0; n < p->n_nodes; n++) {
      if (!state_node[n]) continue;
      if (!tg.filter.matches(P, p->node_taintset[n], node_label_reqs(n))) continue;
      int32_t d;
      if (!node_domain(n, tg.key, &d)) continue;
      tg.register_domain(d);
    }
    for (auto& bp : bound_pods) {
      int cls = bp.first, node = bp.second;
      // TopologyListOptions (topology.go:543-563): nil selector lists every pod of the namespaces
      if (!P.nsset_has(tg.nsset, p->class_namespace[cls])) continue;
      if (tg.selector >= 0 && !P.selector_matches(tg.selector, p->class_labelset[cls])) continue;
      int32_t d;
      if (!node_domain(node, tg.key, &d)) continue;
      if (!tg.filter.matches(P, p->node_taintset[node], node_label_reqs(node))) continue;
      tg.record(d);
    }
  }

  // topology.go:297-322 updateInverseAntiAffinity (required terms only)
  void update_inverse_anti_affinity(int cls, int node /* -1: pending pod */) {
    const kp_problem* p = P.p;
    for (int c = p->class_tsc_off[cls]; c < p->class_tsc_off[cls + 1]; c++) {
      if (p->tsc_type[c] != KP_TOPO_ANTI_AFFINITY) continue;
      if (p->tsc_preferred && p->tsc_preferred[c]) continue;  // required terms only (topology.go:297-322)
      auto g = new_group(cls, c);
      g->inverse = true;
      std::string h = hash_of(*g);
      TopologyGroup* tg;
      auto it = inverse_index.find(h);
      if (it == inverse_index.end()) {
        g->index = (int)inverse_groups.size();
        inverse_index[h] = g->index;
        inverse_groups.push_back(std::move(g));
        tg = inverse_groups.back().get();
      } else {
        tg = inverse_groups[it->second].get();
      }
      if (node >= 0) {
        int32_t d;
        // domains == node.Labels (topology.go:285,316-318): the hostname label only if the node carries it
        if (node_domain(node, tg->key, &d)) tg->record(d);
      }
      tg->owners.insert(cls);
    }
  }

  // topology.go:162-194 Update
  void update(int cls) {
    const kp_problem* p = P.p;
    if (updated.count(cls)) return;  // same class => same groups, AddOwner is idempotent at class level
    updated.insert(cls);
    bool has_anti = false;
    for (int c = p->class_tsc_off[cls]; c < p->class_tsc_off[cls + 1]; c++)
      if (p->tsc_type[c] == KP_TOPO_ANTI_AFFINITY && !(p->tsc_preferred && p->tsc_preferred[c])) has_anti = true;
    if (has_anti) update_inverse_anti_affinity(cls, -1);
    for (int c = p->class_tsc_off[cls]; c < p->class_tsc_off[cls + 1]; c++) {
      auto g = new_group(cls, c);
      std::string h = hash_of(*g);
      TopologyGroup* tg;
      auto it = group_index.find(h);
      if (it == group_index.end()) {
        count_domains(*g);
        g->index = (int)groups.size();
        group_index[h] = g->index;
        groups.push_back(std::move(g));
        tg = groups.back().get();
      } else {
        tg = groups[it->second].get();
      }
      tg->owners.insert(cls);
    }
  }

  // topology.go:68-103 NewTopology: inverse anti-affinities of bound pods, then Update for every pending pod
  void init(const std::vector<int>& pending_classes_in_pod_order) {
    build_domain_groups();
    for (auto& bp : bound_pods) update_inverse_anti_affinity(bp.first, bp.second);
    // the groups of a relaxed pod are created when it is relaxed (Topology.Update in trySchedule, scheduler.go:462)
    for (int cls : pending_classes_in_pod_order) update(cls);
  }

  // topology.go:528-541 getMatchingTopologies
  std::vector<TopologyGroup*> matching(int cls, int taintset, const Requirements& reqs) const {
    std::vector<TopologyGroup*> out;
    for (auto& g : groups)
      if (g->owners.count(cls)) out.push_back(g.get());
    for (auto& g : inverse_groups)
      if (g->counts(P, cls, taintset, reqs)) out.push_back(g.get());
    return out;
  }

  // topology.go:226-248 AddRequirements; returns false on a topologyError
  bool add_requirements(int cls, int taintset, const Requirements& pod_reqs, const Requirements& node_reqs,
                        Requirements* out) const {
    Requirements requirements = node_reqs;
    for (TopologyGroup* tg : matching(cls, taintset, node_reqs)) {
      Requirement pod_domains = pod_reqs.has_key(tg->key) ? pod_reqs.get(tg->key) : exists_requirement(tg->key);
      Requirement node_domains = node_reqs.has_key(tg->key) ? node_reqs.get(tg->key) : exists_requirement(tg->key);
      Requirement domains = tg->get(P, cls, pod_domains, node_domains);
      if (domains.len() == 0) return false;
      requirements.add(P, domains);
    }
    *out = requirements;
    return true;
  }

  // topology.go:197-220 Record
  void record(int cls, int taintset, const Requirements& reqs) {
    for (auto& tg : groups) {
      if (!tg->counts(P, cls, taintset, reqs)) continue;
      Requirement domains = reqs.get(tg->key);
      if (tg->type == KP_TOPO_ANTI_AFFINITY) {
        for (int32_t d : domains.values) tg->record(d);
      } else if (domains.len() == 1) {
        tg->record(domains.values[0]);
      }
    }
    for (auto& tg : inverse_groups)
      if (tg->owners.count(cls))
        for (int32_t d : reqs.get(tg->key).values) tg->record(d);
  }

  // topology.go:251-262 Register
  void register_domain(int key, int32_t domain) {
    for (auto& tg : groups)
      if (tg->key == key) tg->register_domain(domain);
    for (auto& tg : inverse_groups)
      if (tg->key == key) tg->register_domain(domain);
  }
};

}  // namespace orc
