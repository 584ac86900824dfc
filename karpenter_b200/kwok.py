"""KWOK instance-type catalogs: the benchmark input format of the reference.

* `generic_instance_types()` restates kwok/tools/gen_instance_types.go:34-113 (the generator behind the embedded
  kwok/cloudprovider/instance_types.json, 144 types x 8 offerings).
* `aws_instance_types()` loads `data/aws_instance_types.tsv`, a compact table derived from
  kwok/examples/aws_instance_types.json (1724 entries) by tools/make_aws_catalog_fixture.py.
* `new_instance_type()` restates kwok/cloudprovider/helpers.go:131-214 (setDefaultOptions + newInstanceType): the label
  requirements every KWOK instance type carries and the 100m / 10Mi kube-reserved overhead.
"""
from __future__ import annotations

import os
import re
from typing import List

from .model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, OS_LABEL, ZONE_LABEL, InstanceType,
                    NodeSelectorRequirement, Offering)

KWOK_ZONES = ["test-zone-a", "test-zone-b", "test-zone-c", "test-zone-d"]
INSTANCE_SIZE_LABEL = "karpenter.kwok.sh/instance-size"
INSTANCE_FAMILY_LABEL = "karpenter.kwok.sh/instance-family"
INSTANCE_CPU_LABEL = "karpenter.kwok.sh/instance-cpu"
INSTANCE_MEMORY_LABEL = "karpenter.kwok.sh/instance-memory"

_AWS_RE = re.compile(r"^\w+\.(\w+)$")  # helpers.go awsRegexp: "<family>.<size>"
_FAMILY_DELIM = re.compile(r"[.-]")


def _In(key, *values):
    return NodeSelectorRequirement(key, "In", tuple(values))


def new_instance_type(name: str, arch: str, oses: List[str], resources: dict, offerings) -> InstanceType:
    """offerings: list of (capacity_type, zone, price). helpers.go:156-214."""
    cpu, memory = str(resources["cpu"]), str(resources["memory"])
    m = _AWS_RE.match(name)
    size = m.group(1) if m else cpu
    fam = _FAMILY_DELIM.split(name, 1)
    family = fam[0] if len(fam) >= 2 else name[:1]
    res = {"pods": "110"}
    res.update(resources)
    zones, cts = [], []
    for ct, zone, _ in offerings:
        if zone not in zones:
            zones.append(zone)
        if ct not in cts:
            cts.append(ct)
    reqs = [
        _In(INSTANCE_TYPE_LABEL, name), _In(ARCH_LABEL, arch), _In(OS_LABEL, *oses), _In(ZONE_LABEL, *zones),
        _In(CAPACITY_TYPE_LABEL, *cts), _In(INSTANCE_SIZE_LABEL, size), _In(INSTANCE_FAMILY_LABEL, family),
        _In(INSTANCE_CPU_LABEL, cpu), _In(INSTANCE_MEMORY_LABEL, memory),
    ]
    offs = [Offering([_In(CAPACITY_TYPE_LABEL, ct), _In(ZONE_LABEL, zone)], price, True) for ct, zone, price in offerings]
    return InstanceType(name, reqs, offs, res, {"cpu": "100m", "memory": "10Mi"})


def generic_instance_types() -> List[InstanceType]:
    """constructGenericInstanceTypes (gen_instance_types.go:68-111)."""
    out = []
    for cpu in [1, 2, 4, 8, 16, 32, 48, 64, 96, 128, 192, 256]:
        for mem_factor in [2, 4, 8]:
            for os_ in ["linux", "windows"]:
                for arch in ["amd64", "arm64"]:
                    family = {2: "c", 4: "s", 8: "m"}.get(mem_factor, "e")
                    name = f"{family}-{cpu}x-{arch}-{os_}"
                    mem = cpu * mem_factor
                    pods = max(0, min(cpu * 16, 1024))
                    res = {"cpu": str(cpu), "memory": f"{mem}Gi", "pods": str(pods), "ephemeral-storage": "20Gi"}
                    # priceFromResources (gen_instance_types.go:53-66)
                    price = 0.025 * float(cpu) + 0.001 * float(mem * 2**30) / 1e9
                    offs = []
                    for zone in KWOK_ZONES:
                        for ct in ["spot", "on-demand"]:
                            offs.append((ct, zone, price * 0.7 if ct == "spot" else price))
                    out.append(new_instance_type(name, arch, [os_], res, offs))
    return out


_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "aws_instance_types.tsv")
AWS_ZONES = ["us-west-2a", "us-west-2b", "us-west-2c", "us-west-2d"]


def aws_instance_types(n: int = 1724) -> List[InstanceType]:
    """First `n` entries of the reference's kwok/examples/aws_instance_types.json (via the committed TSV)."""
    out = []
    with open(_DATA) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            name, arch, os_, cpu, mem, pods, eph, od, spot = line.rstrip("\n").split("\t")
            offs = []
            for zone in AWS_ZONES:
                offs.append(("spot", zone, float(spot)))
                offs.append(("on-demand", zone, float(od)))
            res = {"cpu": cpu, "memory": mem, "pods": pods, "ephemeral-storage": eph}
            out.append(new_instance_type(name, arch, [os_], res, offs))
            if len(out) >= n:
                break
    return out
