"""The reference's fake cloud provider catalog (test fixture generator), restated so that the reference's own scheduling
test scenarios can be replayed against the oracle and the CUDA path.

    fake.NewInstanceType defaults            pkg/cloudprovider/fake/instancetype.go:50-154
    default GetInstanceTypes() list          pkg/cloudprovider/fake/cloudprovider.go:221-272
    PriceFromResources                       pkg/cloudprovider/fake/instancetype.go:223-236
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from .model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, OS_LABEL, WELL_KNOWN_LABELS, ZONE_LABEL,
                    InstanceType, NodeSelectorRequirement, Offering, parse_quantity)

LABEL_INSTANCE_SIZE = "size"
EXOTIC_INSTANCE_LABEL = "special"
INTEGER_INSTANCE_LABEL = "integer"
GPU_VENDOR_A = "fake.com/vendor-a"
GPU_VENDOR_B = "fake.com/vendor-b"
# the fake provider registers its labels as well-known on import (instancetype.go:41-47)
WELL_KNOWN_LABELS.update({LABEL_INSTANCE_SIZE, EXOTIC_INSTANCE_LABEL, INTEGER_INSTANCE_LABEL})


def price_from_resources(resources: Dict[str, object]) -> float:
    price = 0.0
    for k, v in resources.items():
        if k == "cpu":
            price += 0.1 * float(parse_quantity(v))
        elif k == "memory":
            price += 0.1 * float(parse_quantity(v)) / 1e9
        elif k in (GPU_VENDOR_A, GPU_VENDOR_B):
            price += 1.0
    return price


def _In(key, *values):
    return NodeSelectorRequirement(key, "In", tuple(values))


def new_instance_type(name: str, resources: Optional[Dict[str, object]] = None, architecture: str = "amd64",
                      operating_systems: Sequence[str] = ("linux", "windows", "darwin"),
                      offerings: Optional[List[Offering]] = None) -> InstanceType:
    res = dict(resources or {})
    res.setdefault("cpu", "4")
    res.setdefault("memory", "4Gi")
    res.setdefault("pods", "5")
    if offerings is None:
        price = price_from_resources(res)
        offerings = [Offering([_In(CAPACITY_TYPE_LABEL, ct), _In(ZONE_LABEL, z)], price, True)
                     for ct, z in (("spot", "test-zone-1"), ("spot", "test-zone-2"), ("on-demand", "test-zone-1"),
                                   ("on-demand", "test-zone-2"), ("on-demand", "test-zone-3"))]
    zones, cts = [], []
    for o in offerings:
        if not o.available:
            continue
        for r in o.requirements:
            if r.key == ZONE_LABEL and r.values[0] not in zones:
                zones.append(r.values[0])
            if r.key == CAPACITY_TYPE_LABEL and r.values[0] not in cts:
                cts.append(r.values[0])
    cpu = parse_quantity(res["cpu"])
    large = cpu > 4 and parse_quantity(res["memory"]) > parse_quantity("8Gi")
    reqs = [_In(INSTANCE_TYPE_LABEL, name), _In(ARCH_LABEL, architecture), _In(OS_LABEL, *sorted(operating_systems)),
            _In(ZONE_LABEL, *zones), _In(CAPACITY_TYPE_LABEL, *cts),
            _In(LABEL_INSTANCE_SIZE, "large" if large else "small"),
            _In(INTEGER_INSTANCE_LABEL, str(int(cpu)))]
    reqs.append(_In(EXOTIC_INSTANCE_LABEL, "optional") if large
                else NodeSelectorRequirement(EXOTIC_INSTANCE_LABEL, "DoesNotExist"))
    return InstanceType(name, reqs, offerings, res, {"cpu": "100m", "memory": "10Mi"})


def default_instance_types() -> List[InstanceType]:
    return [
        new_instance_type("default-instance-type"),
        new_instance_type("small-instance-type", {"cpu": "2", "memory": "2Gi"}),
        new_instance_type("gpu-vendor-instance-type", {GPU_VENDOR_A: "2"}),
        new_instance_type("gpu-vendor-b-instance-type", {GPU_VENDOR_B: "2"}),
        new_instance_type("arm-instance-type", {"cpu": "16", "memory": "128Gi"}, architecture="arm64",
                          operating_systems=("ios", "linux", "windows", "darwin")),
        new_instance_type("single-pod-instance-type", {"pods": "1"}),
    ]


def instance_types(total: int) -> List[InstanceType]:
    """fake.InstanceTypes (fake/instancetype.go:200-213): fake-it-<i> with i+1 cpu, 2(i+1) Gi, 10(i+1) pods."""
    return [new_instance_type(f"fake-it-{i}", {"cpu": str(i + 1), "memory": f"{(i + 1) * 2}Gi", "pods": str((i + 1) * 10)})
            for i in range(total)]
