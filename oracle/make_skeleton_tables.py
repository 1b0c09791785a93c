"""TEST/BUILD INFRASTRUCTURE — dumps the reference's skeleton trees (Pose2Sim/skeletons.py, module-level
anytree Nodes) as plain DATA: for every model the pre-order list of [name, id] (children in declaration
order, what RenderTree yields, triangulation.py:735-736).  Run in the build container only.
The result, pose2sim_b200/skeleton_tables.json, is data reused by the drop-in, not code."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ns = ref_shim.load_reference()
from anytree import RenderTree  # noqa: E402  (the shim's minimal anytree)

out = {}
for name in dir(ns.skeletons):
    obj = getattr(ns.skeletons, name)
    if name.isupper() and hasattr(obj, "children") and hasattr(obj, "name"):
        out[name] = [[node.name, getattr(node, "id", None)] for _, _, node in RenderTree(obj)]
path = os.path.join(os.path.dirname(HERE), "pose2sim_b200", "skeleton_tables.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
print({k: len(v) for k, v in out.items()})
