// TEST INFRASTRUCTURE -- CPU oracle. Not part of the product: only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may build, load or call anything under oracle/.
//
// orc_requirement.hpp: restatement of the node-selector set algebra of the reference
//   pkg/scheduling/requirement.go  (Requirement)   and   pkg/scheduling/requirements.go  (Requirements).
// Values are interned integers (local to their label key); the integer reading of a value that Go obtains with
// strconv.Atoi (requirement.go:326-342) comes from an IntTable supplied by the caller.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace orc {

enum Operator { OpIn = 0, OpNotIn = 1, OpExists = 2, OpDoesNotExist = 3, OpGt = 4, OpLt = 5, OpGte = 6, OpLte = 7 };

// value -> integer lookup (strconv.Atoi); returns false when the value is not an integer
struct IntTable {
  virtual ~IntTable() {}
  virtual bool atoi(int key, int32_t value, int64_t* out) const = 0;
};

struct Requirement {
  int key = 0;
  bool complement = false;
  std::vector<int32_t> values;  // sorted, unique
  bool has_gte = false, has_lte = false, has_min = false;
  int64_t gte = 0, lte = 0;
  int32_t min_values = 0;

  bool has_value(int32_t v) const { return std::binary_search(values.begin(), values.end(), v); }
  void insert(int32_t v) {
    auto it = std::lower_bound(values.begin(), values.end(), v);
    if (it == values.end() || *it != v) values.insert(it, v);
  }
  // requirement.go:295-300 -- Len(); complement sets are "infinite minus excluded"
  int64_t len() const { return complement ? INT64_MAX - (int64_t)values.size() : (int64_t)values.size(); }
  // requirement.go:282-293
  Operator op() const {
    if (complement) return len() < INT64_MAX ? OpNotIn : OpExists;
    return len() > 0 ? OpIn : OpDoesNotExist;
  }
  bool operator==(const Requirement& o) const {
    return key == o.key && complement == o.complement && values == o.values && has_gte == o.has_gte &&
           has_lte == o.has_lte && has_min == o.has_min && (!has_gte || gte == o.gte) && (!has_lte || lte == o.lte) &&
           (!has_min || min_values == o.min_values);
  }
};

// requirement.go:326-342 withinBounds
inline bool within_bounds(const IntTable& t, int key, int32_t v, bool has_gte, int64_t gte, bool has_lte, int64_t lte) {
  if (!has_gte && !has_lte) return true;
  int64_t x;
  if (!t.atoi(key, v, &x)) return false;  // with bounds set, non-integer values are invalid
  if (has_gte && x < gte) return false;
  if (has_lte && x > lte) return false;
  return true;
}

// requirement.go:48-102 NewRequirementWithFlexibility (key normalisation is the caller's job: labels.go:117-123).
// `operand` is Atoi(values[0]) for Gt/Lt/Gte/Lte.
inline Requirement new_requirement(int key, Operator op, const std::vector<int32_t>& vals, int64_t operand = 0,
                                   bool has_min = false, int32_t min_values = 0) {
  Requirement r;
  r.key = key;
  r.has_min = has_min;
  r.min_values = min_values;
  if (op == OpIn) {
    for (auto v : vals) r.insert(v);
    r.complement = false;
    return r;
  }
  r.complement = true;
  if (op == OpDoesNotExist) r.complement = false;
  if (op == OpNotIn)
    for (auto v : vals) r.insert(v);
  if (op == OpGt) {
    if (operand == INT64_MAX) {  // Gt MaxInt matches nothing
      Requirement d;
      d.key = key;
      return d;  // DoesNotExist, no minValues (NewRequirement)
    }
    r.has_gte = true;
    r.gte = operand + 1;
  }
  if (op == OpLt) {
    r.has_lte = true;
    r.lte = operand - 1;
  }
  if (op == OpGte) {
    r.has_gte = true;
    r.gte = operand;
  }
  if (op == OpLte) {
    r.has_lte = true;
    r.lte = operand;
  }
  return r;
}

// requirement.go:173-206 Intersection
inline Requirement intersection(const IntTable& t, const Requirement& a, const Requirement& b) {
  Requirement out;
  out.key = a.key;
  bool complement = a.complement && b.complement;
  // maxIntPtr / minIntPtr (requirement.go:344-368)
  bool has_gte = a.has_gte || b.has_gte;
  int64_t gte = a.has_gte && b.has_gte ? std::max(a.gte, b.gte) : (a.has_gte ? a.gte : b.gte);
  bool has_lte = a.has_lte || b.has_lte;
  int64_t lte = a.has_lte && b.has_lte ? std::min(a.lte, b.lte) : (a.has_lte ? a.lte : b.lte);
  bool has_min = a.has_min || b.has_min;
  int32_t mv = a.has_min && b.has_min ? std::max(a.min_values, b.min_values) : (a.has_min ? a.min_values : b.min_values);
  if (has_gte && has_lte && gte > lte) {  // DoesNotExist carrying minValues
    out.has_min = has_min;
    out.min_values = mv;
    return out;
  }
  std::vector<int32_t> vals;
  if (a.complement && b.complement) {
    std::set_union(a.values.begin(), a.values.end(), b.values.begin(), b.values.end(), std::back_inserter(vals));
  } else if (a.complement && !b.complement) {
    std::set_difference(b.values.begin(), b.values.end(), a.values.begin(), a.values.end(), std::back_inserter(vals));
  } else if (!a.complement && b.complement) {
    std::set_difference(a.values.begin(), a.values.end(), b.values.begin(), b.values.end(), std::back_inserter(vals));
  } else {
    std::set_intersection(a.values.begin(), a.values.end(), b.values.begin(), b.values.end(), std::back_inserter(vals));
  }
  for (auto v : vals)
    if (within_bounds(t, a.key, v, has_gte, gte, has_lte, lte)) out.values.push_back(v);
  out.complement = complement;
  if (complement) {  // bounds survive only on complement results
    out.has_gte = has_gte;
    out.gte = gte;
    out.has_lte = has_lte;
    out.lte = lte;
  }
  out.has_min = has_min;
  out.min_values = mv;
  return out;
}

// requirement.go:212-246 HasIntersection
inline bool has_intersection(const IntTable& t, const Requirement& a, const Requirement& b) {
  bool has_gte = a.has_gte || b.has_gte;
  int64_t gte = a.has_gte && b.has_gte ? std::max(a.gte, b.gte) : (a.has_gte ? a.gte : b.gte);
  bool has_lte = a.has_lte || b.has_lte;
  int64_t lte = a.has_lte && b.has_lte ? std::min(a.lte, b.lte) : (a.has_lte ? a.lte : b.lte);
  if (has_gte && has_lte && gte > lte) return false;
  if (a.complement && b.complement) return true;
  if (a.complement && !b.complement) {
    for (auto v : b.values)
      if (!a.has_value(v) && within_bounds(t, a.key, v, has_gte, gte, has_lte, lte)) return true;
    return false;
  }
  if (!a.complement && b.complement) {
    for (auto v : a.values)
      if (!b.has_value(v) && within_bounds(t, a.key, v, has_gte, gte, has_lte, lte)) return true;
    return false;
  }
  for (auto v : a.values)
    if (b.has_value(v) && within_bounds(t, a.key, v, has_gte, gte, has_lte, lte)) return true;
  return false;
}

// requirement.go:267-272 Has
inline bool has(const IntTable& t, const Requirement& r, int32_t v) {
  if (r.complement) return !r.has_value(v) && within_bounds(t, r.key, v, r.has_gte, r.gte, r.has_lte, r.lte);
  return r.has_value(v) && within_bounds(t, r.key, v, r.has_gte, r.gte, r.has_lte, r.lte);
}

inline Requirement exists_requirement(int key) {
  Requirement r;
  r.key = key;
  r.complement = true;
  return r;
}

// requirements.go:36 -- map[string]*Requirement. std::map gives the canonical (ascending key id) iteration order that
// replaces Go's randomised map order.
struct Requirements {
  std::map<int, Requirement> m;

  bool has_key(int key) const { return m.count(key) != 0; }
  // requirements.go:160-166 Get: undefined == Exists
  Requirement get(int key) const {
    auto it = m.find(key);
    if (it == m.end()) return exists_requirement(key);
    return it->second;
  }
  // requirements.go:133-140 Add: intersect on key collision
  void add(const IntTable& t, const Requirement& r) {
    auto it = m.find(r.key);
    if (it != m.end()) {
      it->second = intersection(t, r, it->second);
    } else {
      m.emplace(r.key, r);
    }
  }
  void add_all(const IntTable& t, const Requirements& o) {
    for (auto& kv : o.m) add(t, kv.second);
  }
  bool has_min_values() const {
    for (auto& kv : m)
      if (kv.second.has_min) return true;
    return false;
  }
};

struct KeyPolicy {
  virtual ~KeyPolicy() {}
  virtual bool well_known(int key) const = 0;  // v1.WellKnownLabels (labels.go:69-78)
};

// requirements.go:254-274 Intersects. Returns true when there is NO error.
inline bool intersects(const IntTable& t, const Requirements& r, const Requirements& incoming) {
  const Requirements* small = &r;
  const Requirements* large = &incoming;
  if (small->m.size() > large->m.size()) std::swap(small, large);
  bool ok = true;
  for (auto& kv : small->m) {
    int key = kv.first;
    if (!large->has_key(key)) continue;
    const Requirement& existing = r.m.at(key);
    const Requirement& inc = incoming.m.at(key);
    if (!has_intersection(t, existing, inc)) {
      Operator io = inc.op();
      if (io == OpNotIn || io == OpDoesNotExist) {
        Operator eo = existing.op();
        if (eo == OpNotIn || eo == OpDoesNotExist) continue;
      }
      ok = false;
    }
  }
  return ok;
}

// requirements.go:181-197 Compatible. allow_undefined_well_known == the AllowUndefinedWellKnownLabels option.
// Returns true when compatible (error == nil).
inline bool compatible(const IntTable& t, const KeyPolicy& kp, const Requirements& r, const Requirements& incoming,
                       bool allow_undefined_well_known) {
  for (auto& kv : incoming.m) {
    int key = kv.first;
    if (allow_undefined_well_known && kp.well_known(key)) continue;
    Operator o = kv.second.op();
    if (r.has_key(key) || o == OpNotIn || o == OpDoesNotExist) continue;
    return false;  // label does not have known values
  }
  return intersects(t, r, incoming);
}

}  // namespace orc
