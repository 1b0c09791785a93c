"""Keypoint tables of the pose models, as plain data.

The reference keeps its skeletons as module-level `anytree` trees (Pose2Sim/skeletons.py:50-981) and
only ever uses their PRE-ORDER traversal: ids/names with `id != None` define the keypoint (and TRC
marker) order (triangulation.py:735-736), the tracked keypoint is looked up by name
(personAssociation.py:749).  `skeleton_tables.json` holds that traversal for every shipped model
(dumped from the reference by oracle/make_skeleton_tables.py); a custom model given in Config.toml
as a nested dict (`[pose.CUSTOM]`, triangulation.py:726-730) is traversed here the same way.
"""
import json
import logging
import os

_TABLES = None

# triangulation.py:717-724 / personAssociation.py:692-699
ALIASES = {"BODY_WITH_FEET": "HALPE_26", "WHOLE_BODY_WRIST": "COCO_133_WRIST", "WHOLE_BODY": "COCO_133",
           "BODY": "COCO_17", "HAND": "HAND_21", "FACE": "FACE_106", "ANIMAL": "ANIMAL2D_17"}


def _tables():
    global _TABLES
    if _TABLES is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "skeleton_tables.json")) as f:
            _TABLES = json.load(f)
    return _TABLES


# id given to a node of a Config.toml model whose id cannot index a keypoint list (see _preorder_dict): larger than any
# pose_keypoints_2d, so every lookup misses and the marker is NaN in every frame
UNREADABLE_ID = 100_000_000


def _preorder_dict(node, out, root=True):
    """Pre-order [(name, id)] of a `[pose.CUSTOM]` table.  The reference turns the string 'None' into None for the ROOT
    only (triangulation.py:727-729); any other node keeps a string id, passes the `id != None` filter (:735) and becomes a
    marker whose lookup `keypoints[id * 3]` always fails into the bare `except` (:629-644): a marker that is NaN in every
    frame.  Such nodes get UNREADABLE_ID here — same marker list, same NaN columns."""
    nid = node.get("id")
    if isinstance(nid, str):
        nid = None if (root and nid == "None") else UNREADABLE_ID
    out.append([node.get("name"), nid])
    for child in node.get("children", []) or []:
        _preorder_dict(child, out, root=False)
    return out


def model_nodes(pose_model, config_dict=None):
    """Pre-order [(name, id)] of the model; raises NameError like the reference when unknown."""
    name = ALIASES.get(str(pose_model).upper(), pose_model)
    tables = _tables()
    if name in tables:
        return [tuple(n) for n in tables[name]]
    custom = ((config_dict or {}).get("pose") or {}).get(pose_model)
    if isinstance(custom, dict):
        return [tuple(n) for n in _preorder_dict(custom, [])]
    raise NameError(f"{pose_model} not found in skeletons.py nor in Config.toml")


def keypoints(pose_model, config_dict=None):
    """(ids, names) of the keypoints with an id, in pre-order (triangulation.py:735-736)."""
    nodes = model_nodes(pose_model, config_dict)
    ids = [i for _, i in nodes if i is not None]
    names = [n for n, i in nodes if i is not None]
    return ids, names


def swapped_indices(names):
    """`keypoints_idx_swapped` (triangulation.py:741-749): the index of each keypoint's left/right partner — an
    initial 'R' <-> 'L', then a leading 'right' <-> 'left'; a partner name that does not exist disables the swap
    for ALL keypoints (the reference's bare `except`), with its warning."""
    names = list(names)
    try:
        sw = ["L" + n[1:] if n.startswith("R") else "R" + n[1:] if n.startswith("L") else n for n in names]
        sw = [n.replace("right", "left") if n.startswith("right") else n.replace("left", "right") if n.startswith("left") else n
              for n in sw]
        return [names.index(n) for n in sw]
    except ValueError:
        logging.warning("No left/right swap was performed.")
        return list(range(len(names)))


def tracked_keypoint_id(pose_model, tracked_keypoint, config_dict=None):
    """personAssociation.py:747-753: id of the first node named `tracked_keypoint`; falls back to id 0
    (with the name of the node carrying id 0) when missing or falsy.  Returns (id, fallback_name|None)."""
    nodes = model_nodes(pose_model, config_dict)
    for n, i in nodes:
        if n == tracked_keypoint:
            if i:
                return i, None
            break
    fallback = next((n for n, i in nodes if i == 0), None)
    return 0, fallback
