"""TEST INFRASTRUCTURE — not product code.

Import shim that lets the UNMODIFIED reference (`/root/reference/Pose2Sim`) be imported in the
build container so that it can serve as the live parity authority and generate the golden vectors
committed under `tests/golden/` (see `oracle/make_golden.py`).

The reference cannot be imported as-is here because:
  * every module calls `importlib.metadata.version('pose2sim')` at import time
    (`Pose2Sim/common.py:43-44`, `triangulation.py:69-70`, `personAssociation.py:59-60`, `skeletons.py:40-41`);
  * `common.py:23,29-33` imports `c3d`, `tkinter`, `matplotlib`, `PyQt5` at module level;
  * `skeletons.py:32`, `triangulation.py:54-55`, `personAssociation.py:42-43` need `anytree`.

Nothing here is used on the GPU box: `/root/reference` does not exist there, and only
`oracle/make_golden.py` and the container-only cross-checks in `tests/` call `load_reference()`.
"""
import importlib
import importlib.metadata
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("P2S_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Pose2Sim"))


class _Dummy:
    """Attribute sink: any attribute is another dummy class; calling returns a dummy."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Dummy,), {})


def _stub(name):
    if name not in sys.modules:
        m = _StubModule(name)
        m.__path__ = []  # make it a package so that sub-imports resolve
        sys.modules[name] = m
    return sys.modules[name]


# ---------------------------------------------------------------------------------------------
# minimal `anytree` (Node, RenderTree, PreOrderIter, importer.DictImporter): pre-order traversal
# with children in declaration order is all the reference relies on (triangulation.py:735-736).
# ---------------------------------------------------------------------------------------------
class Node:
    def __init__(self, name, parent=None, children=None, **attrs):
        self.name = name
        self.parent = None
        self.children = []
        for k, v in attrs.items():
            setattr(self, k, v)
        if parent is not None:
            self.parent = parent
            parent.children.append(self)
        if children:
            for c in children:
                c.parent = self
                self.children.append(c)

    @property
    def is_leaf(self):
        return len(self.children) == 0

    @property
    def root(self):
        n = self
        while n.parent is not None:
            n = n.parent
        return n

    @property
    def path(self):
        p, n = [], self
        while n is not None:
            p.append(n)
            n = n.parent
        return tuple(reversed(p))

    @property
    def descendants(self):
        return tuple(PreOrderIter(self))[1:]

    def __repr__(self):
        return f"Node({self.name!r})"


def PreOrderIter(node, filter_=None):
    stack = [node]
    while stack:
        n = stack.pop()
        if filter_ is None or filter_(n):
            yield n
        stack.extend(reversed(n.children))


def RenderTree(node):
    for n in PreOrderIter(node):
        depth = len(n.path) - 1
        yield ("    " * depth, "    " * depth, n)


class DictImporter:
    def __init__(self, nodecls=Node):
        self.nodecls = nodecls

    def import_(self, data):
        return self._imp(dict(data), None)

    def _imp(self, data, parent):
        children = data.pop("children", [])
        name = data.pop("name", None)
        node = self.nodecls(name, parent=parent, **data)
        for c in children:
            self._imp(dict(c), node)
        return node


def _install_anytree():
    try:
        import anytree  # noqa: F401  (a real one wins if present)
        return
    except ImportError:
        pass
    m = types.ModuleType("anytree")
    m.Node, m.RenderTree, m.PreOrderIter = Node, RenderTree, PreOrderIter
    imp = types.ModuleType("anytree.importer")
    imp.DictImporter = DictImporter
    m.importer = imp
    m.__path__ = []
    sys.modules["anytree"] = m
    sys.modules["anytree.importer"] = imp


_loaded = None


def load_reference():
    """Return a namespace with the reference modules `common`, `triangulation`, `personAssociation`,
    `skeletons` imported unmodified from REFERENCE_ROOT."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (only in the build container)")
    for name in ("c3d", "tkinter", "matplotlib", "matplotlib.pyplot", "matplotlib.backends",
                 "matplotlib.backends.backend_qt5agg", "PyQt5", "PyQt5.QtWidgets", "PyQt5.QtCore",
                 "PyQt5.QtGui"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    _install_anytree()
    real_version = importlib.metadata.version

    def version(dist):
        if dist.lower() == "pose2sim":
            return "0.10.0+reference"
        return real_version(dist)

    importlib.metadata.version = version
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.common = importlib.import_module("Pose2Sim.common")
    ns.skeletons = importlib.import_module("Pose2Sim.skeletons")
    ns.triangulation = importlib.import_module("Pose2Sim.triangulation")
    ns.personAssociation = importlib.import_module("Pose2Sim.personAssociation")
    _loaded = ns
    return ns
