"""URDF -> Isaac Gym body / DOF ordering and joint limits (oracle; TEST INFRASTRUCTURE ONLY).

Isaac Gym orders rigid bodies and DOFs by a depth-first traversal of the kinematic tree with the children of
every link sorted by JOINT NAME.  Applied to the reference's
``resources/assets/bez/model/soccerbot_stl.urdf`` this reproduces every index the reference hard-codes
(``bez_isaacgym/tasks/kick_env.py:23-41,175-177,188-196``), which is what pins the tensor layout.  Used by
``oracle/fake_isaacgym.py`` (so the unmodified reference runs with limits parsed from its own URDF) and by
``tests/test_model_constants.py`` (to check the product's hard-coded constants).
"""
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import List


@dataclass
class Layout:
    bodies: List[str]
    dof_names: List[str]
    lower: List[float]
    upper: List[float]
    num_joints: int


def parse(urdf_path: str) -> Layout:
    root = ET.parse(urdf_path).getroot()
    joints = []
    children_of = {}
    child_links = set()
    for j in root.findall("joint"):
        name, kind = j.get("name"), j.get("type")
        parent, child = j.find("parent").get("link"), j.find("child").get("link")
        lim = j.find("limit")
        lo = float(lim.get("lower", 0.0)) if lim is not None else 0.0
        hi = float(lim.get("upper", 0.0)) if lim is not None else 0.0
        rec = dict(name=name, kind=kind, parent=parent, child=child, lower=lo, upper=hi)
        joints.append(rec)
        children_of.setdefault(parent, []).append(rec)
        child_links.add(child)
    links = [l.get("name") for l in root.findall("link")]
    base = [l for l in links if l not in child_links]
    assert len(base) == 1, f"expected one root link, got {base}"
    bodies, dofs, lower, upper = [], [], [], []

    def visit(link):
        bodies.append(link.lstrip("/"))
        for rec in sorted(children_of.get(link, []), key=lambda r: r["name"]):
            if rec["kind"] in ("revolute", "prismatic", "continuous"):
                dofs.append(rec["name"])
                lower.append(rec["lower"])
                upper.append(rec["upper"])
            visit(rec["child"])

    visit(base[0])
    return Layout(bodies, dofs, lower, upper, len(joints))
