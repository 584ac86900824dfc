"""Host-side mirror of the reference's object model for the provisioning hot path.

These are plain data holders named after the Go types the reference's callers hand to
``scheduling.NewScheduler`` / ``Scheduler.Solve`` (pkg/controllers/provisioning/scheduling/scheduler.go:116-129,381)
and the plugin types of pkg/cloudprovider/types.go:122-138,372-379.  They carry strings; `encode.py` interns them into
the flat ``kp_problem`` of include/karpsolve.h.  No scheduling logic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

# pkg/apis/v1/labels.go:69-78,117-123 and k8s.io/api well-known label names
NODEPOOL_LABEL = "karpenter.sh/nodepool"
CAPACITY_TYPE_LABEL = "karpenter.sh/capacity-type"
ZONE_LABEL = "topology.kubernetes.io/zone"
REGION_LABEL = "topology.kubernetes.io/region"
INSTANCE_TYPE_LABEL = "node.kubernetes.io/instance-type"
ARCH_LABEL = "kubernetes.io/arch"
OS_LABEL = "kubernetes.io/os"
WINDOWS_BUILD_LABEL = "node.kubernetes.io/windows-build"
HOSTNAME_LABEL = "kubernetes.io/hostname"

WELL_KNOWN_LABELS = {
    NODEPOOL_LABEL, ZONE_LABEL, REGION_LABEL, INSTANCE_TYPE_LABEL, ARCH_LABEL, OS_LABEL, CAPACITY_TYPE_LABEL,
    WINDOWS_BUILD_LABEL,
}
NORMALIZED_LABELS = {
    "failure-domain.beta.kubernetes.io/zone": ZONE_LABEL,
    "beta.kubernetes.io/arch": ARCH_LABEL,
    "beta.kubernetes.io/os": OS_LABEL,
    "beta.kubernetes.io/instance-type": INSTANCE_TYPE_LABEL,
    "failure-domain.beta.kubernetes.io/region": REGION_LABEL,
}

CAPACITY_TYPE_SPOT = "spot"
CAPACITY_TYPE_ON_DEMAND = "on-demand"
CAPACITY_TYPE_RESERVED = "reserved"

_SUFFIX = {
    "n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "": Fraction(1),
    "k": Fraction(10**3), "M": Fraction(10**6), "G": Fraction(10**9), "T": Fraction(10**12),
    "P": Fraction(10**15), "E": Fraction(10**18),
    "Ki": Fraction(2**10), "Mi": Fraction(2**20), "Gi": Fraction(2**30), "Ti": Fraction(2**40),
    "Pi": Fraction(2**50), "Ei": Fraction(2**60),
}


def parse_quantity(q) -> Fraction:
    """resource.MustParse (k8s.io/apimachinery/pkg/api/resource) for the decimal/binary-SI forms the reference uses."""
    if isinstance(q, (int, Fraction)):
        return Fraction(q)
    s = str(q).strip()
    for suf in sorted(_SUFFIX, key=len, reverse=True):
        if suf and s.endswith(suf):
            return Fraction(s[: -len(suf)]) * _SUFFIX[suf]
    if "e" in s or "E" in s:
        mant, exp = s.replace("E", "e").split("e")
        return Fraction(mant) * Fraction(10) ** int(exp)
    return Fraction(s)


_UNITS_CACHE: Dict[Tuple[str, object], int] = {}


def quantity_units(name: str, q) -> int:
    """Exact integer in the resource's unit: milli for cpu, base units (bytes / count) otherwise.  Memoised: a pass over
    10^5 pods parses the same few quantity strings over and over (the reference caches PodData per pod for the same reason,
    scheduler.go:471-491)."""
    if isinstance(q, (str, int)):
        hit = _UNITS_CACHE.get((name, q))
        if hit is not None:
            return hit
        v = _quantity_units(name, q)
        if len(_UNITS_CACHE) < 65536:
            _UNITS_CACHE[(name, q)] = v
        return v
    return _quantity_units(name, q)


def _quantity_units(name: str, q) -> int:
    f = parse_quantity(q)
    if name == "cpu":
        f = f * 1000
    if f.denominator != 1:
        raise ValueError(f"quantity {q!r} of {name} is not integral in the solver's unit")
    return int(f)


def _q(name, v):
    return quantity_units(name, v)


def host_port_key(hp) -> Tuple[str, int, str]:
    """(ip, port, protocol) in canonical form: GetHostPorts' defaults (hostportusage.go:93-118) and one spelling for the
    unspecified address (net.IP.IsUnspecified: 0.0.0.0 and ::)."""
    ip, port, proto = hp
    ip = "" if ip in ("", "0.0.0.0", "::") else ip
    return (ip, int(port), proto or "TCP")


def host_ports_match(a, b) -> bool:
    """HostPort.Matches (hostportusage.go:50-62): same protocol and port, IPs equal or one of them unspecified."""
    a, b = host_port_key(a), host_port_key(b)
    return a[2] == b[2] and a[1] == b[1] and (a[0] == b[0] or a[0] == "" or b[0] == "")


def _ceiling_side(main: Dict[str, object], inits, side: str, overhead: Dict[str, object], pod_level: Dict[str, object]) -> Dict[str, int]:
    """One side (requests or limits) of resources.Ceiling == component-helpers resource.PodRequests / PodLimits
    (pkg/utils/resources/resources.go:113-118; k8s.io/component-helpers v0.35 resource/helpers.go):
      total  = sum(containers) + sum(sidecars)
      init_i = requests(init_i) + sum(sidecars declared BEFORE init_i)      (a sidecar counts itself plus the earlier ones)
      result = max(total, max_i init_i) per resource, pod-level resources override the names they carry, + overhead."""
    total = {k: _q(k, v) for k, v in main.items()}
    sidecars: Dict[str, int] = {}
    peak: Dict[str, int] = {}
    for c in inits:
        cur = {k: _q(k, v) for k, v in getattr(c, side).items()}
        if c.restart_always:
            for k, v in cur.items():
                total[k] = total.get(k, 0) + v
                sidecars[k] = sidecars.get(k, 0) + v
            cur = dict(sidecars)
        else:
            for k, v in sidecars.items():
                cur[k] = cur.get(k, 0) + v
        for k, v in cur.items():
            if v > peak.get(k, 0) or k not in peak:
                peak[k] = max(v, peak.get(k, v))
    for k, v in peak.items():
        if v > total.get(k, 0) or k not in total:
            total[k] = max(v, total.get(k, v))
    for k, v in pod_level.items():  # PodLevelResources: cpu / memory / hugepages replace the aggregate
        if k in ("cpu", "memory") or k.startswith("hugepages-"):
            total[k] = _q(k, v)
    if overhead:
        for k, v in overhead.items():
            total[k] = total.get(k, 0) + _q(k, v)
    return total


def ceiling(pod) -> Tuple[Dict[str, int], Dict[str, int]]:
    """resources.Ceiling(pod) -> (requests, limits) in the solver's exact integer units (milli-cpu, bytes, counts).  Overhead
    is added to a limit only where a limit exists (PodLimits)."""
    req = _ceiling_side(pod.requests, pod.init_containers, "requests", pod.overhead, pod.pod_level_requests)
    lim = _ceiling_side(pod.limits, pod.init_containers, "limits", {}, pod.pod_level_limits)
    for k, v in pod.overhead.items():
        if k in lim:
            lim[k] += _q(k, v)
    return req, lim


def effective_requests(pod) -> Dict[str, int]:
    """PodData.Requests without the pods: 1 entry (scheduler.go:471-491, resources.go:30-39): what the encoder puts into the
    class row.  A pod without init containers / overhead / pod-level resources keeps its `requests`."""
    if not (pod.init_containers or pod.overhead or pod.pod_level_requests):
        return {k: _q(k, v) for k, v in pod.requests.items()}
    return ceiling(pod)[0]


@dataclass(frozen=True)
class NodeSelectorRequirement:
    """corev1.NodeSelectorRequirement / v1.NodeSelectorRequirementWithMinValues."""
    key: str
    operator: str  # In NotIn Exists DoesNotExist Gt Lt Gte Lte
    values: Tuple[str, ...] = ()
    min_values: Optional[int] = None

    def __post_init__(self):
        object.__setattr__(self, "values", tuple(self.values))


@dataclass(frozen=True)
class Taint:
    key: str
    value: str = ""
    effect: str = "NoSchedule"


@dataclass(frozen=True)
class Toleration:
    key: str = ""
    operator: str = "Equal"  # "" == Equal
    value: str = ""
    effect: str = ""


@dataclass(frozen=True)
class LabelSelector:
    match_labels: Tuple[Tuple[str, str], ...] = ()
    match_expressions: Tuple[Tuple[str, str, Tuple[str, ...]], ...] = ()  # (key, In|NotIn|Exists|DoesNotExist, values)

    @staticmethod
    def of(match_labels: Optional[Dict[str, str]] = None, match_expressions=()):
        ml = tuple(sorted((match_labels or {}).items()))
        me = tuple((k, op, tuple(vs)) for k, op, vs in match_expressions)
        return LabelSelector(ml, me)


@dataclass(frozen=True)
class TopologySpreadConstraint:
    max_skew: int
    topology_key: str
    label_selector: Optional[LabelSelector]
    when_unsatisfiable: str = "DoNotSchedule"
    min_domains: Optional[int] = None
    node_taints_policy: Optional[str] = None    # Honor | Ignore (default Ignore)
    node_affinity_policy: Optional[str] = None  # Honor | Ignore (default Honor)
    match_label_keys: Tuple[str, ...] = ()


@dataclass(frozen=True)
class PodAffinityTerm:
    label_selector: Optional[LabelSelector]
    topology_key: str
    namespaces: Tuple[str, ...] = ()


@dataclass(frozen=True)
class PreferredSchedulingTerm:
    """corev1.PreferredSchedulingTerm: a soft node-affinity term."""
    weight: int
    match_expressions: Tuple[NodeSelectorRequirement, ...] = ()

    def __post_init__(self):
        object.__setattr__(self, "match_expressions", tuple(self.match_expressions))


@dataclass(frozen=True)
class WeightedPodAffinityTerm:
    """corev1.WeightedPodAffinityTerm: a soft pod affinity / anti-affinity term."""
    weight: int
    term: PodAffinityTerm


@dataclass
class Container:
    """corev1.Container as resources.Ceiling reads it: requests / limits, and RestartPolicy == Always for an init container
    that is a sidecar."""
    requests: Dict[str, object] = field(default_factory=dict)
    limits: Dict[str, object] = field(default_factory=dict)
    restart_always: bool = False


@dataclass
class Pod:
    name: str = ""
    namespace: str = "default"
    uid: int = 0  # 128-bit integer; ordered like the canonical UUID string
    labels: Dict[str, str] = field(default_factory=dict)
    requests: Dict[str, object] = field(default_factory=dict)  # resource name -> quantity (str|int)
    node_selector: Dict[str, str] = field(default_factory=dict)
    node_affinity_required: List[List[NodeSelectorRequirement]] = field(default_factory=list)  # OR of terms
    tolerations: List[Toleration] = field(default_factory=list)
    topology_spread_constraints: List[TopologySpreadConstraint] = field(default_factory=list)
    pod_affinity: List[PodAffinityTerm] = field(default_factory=list)       # required terms
    pod_anti_affinity: List[PodAffinityTerm] = field(default_factory=list)  # required terms
    # soft constraints: enforced until Preferences.Relax drops them one at a time (preferences.go:38-57)
    node_affinity_preferred: List[PreferredSchedulingTerm] = field(default_factory=list)
    pod_affinity_preferred: List[WeightedPodAffinityTerm] = field(default_factory=list)
    pod_anti_affinity_preferred: List[WeightedPodAffinityTerm] = field(default_factory=list)
    # volumeReqsByPod[uid] (scheduler.go:127,489): alternatives of topology requirements of the pod's volumes, as
    # VolumeTopology.GetRequirements derives them (volumetopology.go:44-89).  They constrain the node, never the pod's own
    # topology domains (nodeclaim.go:136-176).  One alternative is supported (a bound PV / one storage class).
    volume_requirements: List[List[NodeSelectorRequirement]] = field(default_factory=list)
    creation_timestamp: int = 0
    # disruption cost inputs (utils/disruption/disruption.go:48-70): Spec.Priority and the
    # controller.kubernetes.io/pod-deletion-cost annotation (None: absent)
    priority: Optional[int] = None
    deletion_cost: Optional[str] = None
    # what resources.Ceiling folds into the effective requests besides the containers' sum (`requests` / `limits` above
    # are that sum): init containers in spec order, sidecars among them (restart_always); RuntimeClass overhead; pod-level
    # resources (Spec.Resources), which replace the aggregate for the resources they name
    limits: Dict[str, object] = field(default_factory=dict)
    init_containers: List["Container"] = field(default_factory=list)
    overhead: Dict[str, object] = field(default_factory=dict)
    pod_level_requests: Dict[str, object] = field(default_factory=dict)
    pod_level_limits: Dict[str, object] = field(default_factory=dict)
    # container ports with a hostPort: (hostIP, hostPort, protocol); "" / "0.0.0.0" / "::" are the unspecified address, the
    # protocol defaults to TCP (scheduling.GetHostPorts, hostportusage.go:93-118)
    host_ports: List[Tuple[str, int, str]] = field(default_factory=list)
    # Optional: an opaque key two pods share only if everything the scheduler looks at is identical (the owner's
    # pod-template-hash in practice).  The encoder then interns the spec once per key instead of once per pod -- what the
    # reference's per-pod PodData cache (scheduler.go:471-491) buys it, and the shim's job when 10^6 pods arrive.
    template: Optional[object] = None


# cloudprovider.ReservationIDLabel (pkg/cloudprovider/types.go:49-52) is the provider's to name; this is the fake provider's
# (pkg/test/v1alpha1/labels.go:20), which also registers it as a well-known label (fake/cloudprovider.go:44-48)
RESERVATION_ID_LABEL = "karpenter.test.sh/reservation-id"
WELL_KNOWN_LABELS.add(RESERVATION_ID_LABEL)


@dataclass
class Offering:
    requirements: List[NodeSelectorRequirement]
    price: float
    available: bool = True
    reservation_capacity: int = 0  # Offering.ReservationCapacity (types.go:372-379); the id is the reservation-id requirement


@dataclass
class InstanceType:
    name: str
    requirements: List[NodeSelectorRequirement]
    offerings: List[Offering]
    capacity: Dict[str, object]
    overhead: Dict[str, object] = field(default_factory=dict)  # Overhead.Total()


@dataclass
class NodePool:
    name: str
    weight: int = 0
    requirements: List[NodeSelectorRequirement] = field(default_factory=list)
    labels: Dict[str, str] = field(default_factory=dict)
    taints: List[Taint] = field(default_factory=list)
    limits: Dict[str, object] = field(default_factory=dict)
    node_class: str = "default"
    # spec.template.spec.nodeClassRef group / kind: the NodeClaim template carries the label NodeClassLabelKey(gk) =
    # "<group>/<lower(kind)>" (pkg/apis/v1/labels.go:162-164, nodeclaimtemplate.go:60-65); the defaults are KWOK's
    node_class_group: str = "karpenter.kwok.sh"
    node_class_kind: str = "KWOKNodeClass"

    def node_class_label_key(self) -> str:
        return f"{self.node_class_group}/{self.node_class_kind.lower()}"


@dataclass
class StateNode:
    """state.StateNode as read by NewExistingNode (existingnode.go:40-66) and the disruption Candidate (types.go:74-82)."""
    name: str
    labels: Dict[str, str] = field(default_factory=dict)
    taints: List[Taint] = field(default_factory=list)
    available: Dict[str, object] = field(default_factory=dict)  # Allocatable - pod requests (- remaining daemons)
    capacity: Dict[str, object] = field(default_factory=dict)
    initialized: bool = True
    managed: bool = True
    schedulable: bool = True
    nodepool: Optional[str] = None
    instance_type: Optional[str] = None
    pods: List[Pod] = field(default_factory=list)          # reschedulable pods bound to the node (consolidation)
    running_pods: List[Pod] = field(default_factory=list)  # other bound pods, only counted by the topology
    # NodeClaim lifetime for LifetimeRemaining (utils/disruption/disruption.go:36-46): Spec.ExpireAfter in seconds
    # (None: never) and the NodeClaim's age in seconds
    expire_after_s: Optional[float] = None
    age_s: float = 0.0
    # StateNode.HostPortUsage(): host ports of pods bound to the node that are not listed in `pods` / `running_pods`
    host_ports: List[Tuple[str, int, str]] = field(default_factory=list)


@dataclass
class NodeClaimResult:
    """scheduling.NodeClaim as callers read it (nodeclaim.go:40-60): template, pods, instance type options, requirements."""
    nodepool: str
    pods: List[Pod]
    instance_type_options: List[str]
    requirements: Dict[str, dict]
    requests: Dict[str, int]
    rank: int


@dataclass
class Results:
    """scheduling.Results (scheduler.go:237-241)."""
    new_node_claims: List[NodeClaimResult]
    existing_nodes: Dict[str, List[Pod]]
    pod_errors: Dict[int, str]  # id(pod) -> reason
    raw: object = None
