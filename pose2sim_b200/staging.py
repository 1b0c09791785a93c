"""Host staging: OpenPose-format JSON directories -> packed float32 arrays for ONE batched device call.

The reference re-opens every camera's JSON once per person and rescans the whole file list per frame
(Pose2Sim/triangulation.py:796-803, :607-653; Pose2Sim/personAssociation.py:758-774, :67-99, :260-274).
Here every file is parsed once, but the VALUES and their ORDER follow the reference exactly:

  * camera directories: sub-directories of `pose/` whose name contains 'json', ordered by the last
    number in the name (common.py:568-583, triangulation.py:752-758);
  * the input directory is pose-associated -> pose-sync -> pose for triangulation (:759-771) and
    pose-sync -> pose for association (personAssociation.py:724-731);
  * the file of frame f in camera c is the one whose LAST number equals f (:799), 'none' if absent;
  * keypoint k of person n in camera c is `people[n].pose_keypoints_2d[3*id_k : 3*id_k+3]`, any
    failure (missing file, person, short list) is a NaN triple (:626-644).

Observations are stored as float32 (north_star: float4 SoA staging).  Pose2Sim's own pose stage writes
float32 values (`kp[0].item()`, poseEstimation.py:260), so nothing is lost for its files; when a JSON
holds values that float32 cannot represent, `float32_inexact` counts them and the callers log it.
"""
import fnmatch
import functools
import json
import os
import re

import numpy as np

_LAST_NUMBER = re.compile(r"\d+")


@functools.lru_cache(maxsize=1 << 20)
def _last_int(name):
    """Last run of digits of `name` as an int, None without digits (one regex pass per distinct name: the sort
    keys and the frame table of a 40 k-file trial share it)."""
    nums = _LAST_NUMBER.findall(name)
    return int(nums[-1]) if nums else None


def sort_by_last_number(names):
    """common.py:568-583: strings with a number first, by their last number; the rest alphabetically."""
    def key(s):
        n = _last_int(s)
        return (True, s) if n is None else (False, n)
    return sorted(names, key=key)


def frame_number(name):
    """triangulation.py:799 `int(re.split(r'(\\d+)', j)[-2])`: the last run of digits in the name.
    Raises IndexError for a name without digits, like the reference."""
    n = _last_int(name)
    if n is None:
        raise IndexError("list index out of range")
    return n


class PoseDirs:
    """Directory discovery shared by the two stages."""

    def __init__(self, project_dir):
        self.project_dir = project_dir
        self.pose_dir = os.path.join(project_dir, "pose")
        self.sync_dir = os.path.join(project_dir, "pose-sync")
        self.tracked_dir = os.path.join(project_dir, "pose-associated")

    def camera_dirs(self):
        """triangulation.py:752-758: the probe looks into the first folder in `os.walk` order (unsorted)."""
        try:
            names = next(os.walk(self.pose_dir))[1]
            os.listdir(os.path.join(self.pose_dir, names[0]))[0]
        except Exception:
            raise ValueError(f"No json files found in {self.pose_dir} subdirectories. "
                             f"Make sure you run Pose2Sim.poseEstimation() first.")
        names = sort_by_last_number(names)
        return [k for k in names if "json" in k]

    def camera_dirs_for_association(self):
        """personAssociation.py:713-720 — not the same statement order as the triangulation stage: the walk happens
        outside the `try` (a missing pose folder surfaces as the generator's StopIteration), and the probe looks into the
        first folder AFTER sorting, so another camera's folder may be empty (that camera is then 'none' in every frame)."""
        names = next(os.walk(self.pose_dir))[1]
        try:
            names = sort_by_last_number(names)
            os.listdir(os.path.join(self.pose_dir, names[0]))[0]
        except Exception:
            raise ValueError(f"No json files found in {self.pose_dir} subdirectories. "
                             f"Make sure you run Pose2Sim.poseEstimation() first.")
        return [k for k in names if "json" in k]

    @staticmethod
    def _list(base, cam_dirs):
        return [fnmatch.filter(os.listdir(os.path.join(base, d)), "*.json") for d in cam_dirs]

    def files_for_triangulation(self, cam_dirs):
        """Returns (input_dir, per-camera sorted file names): pose-associated, else pose-sync, else pose."""
        for base in (self.tracked_dir, self.sync_dir, self.pose_dir):
            try:
                names = self._list(base, cam_dirs)
                return base, [sort_by_last_number(n) for n in names]
            except Exception:
                continue
        raise Exception(f"No json files found in {self.pose_dir}, {self.sync_dir}, nor {self.tracked_dir} "
                        f"subdirectories. Make sure you run Pose2Sim.poseEstimation() first.")

    def index_for_triangulation(self, cam_dirs):
        """`files_for_triangulation` through the native lister: (input_dir, NativeIndex)."""
        for base in (self.tracked_dir, self.sync_dir, self.pose_dir):
            try:
                return base, NativeIndex(base, cam_dirs)
            except OSError:
                continue
        raise Exception(f"No json files found in {self.pose_dir}, {self.sync_dir}, nor {self.tracked_dir} "
                        f"subdirectories. Make sure you run Pose2Sim.poseEstimation() first.")

    def files_for_association(self, cam_dirs):
        """pose-sync, else pose (personAssociation.py:724-731).  N.B. the reference LISTS pose-sync but
        then READS from pose/ because of an `os.path.exist` typo (:762-766); both are kept."""
        for base in (self.sync_dir, self.pose_dir):
            try:
                names = self._list(base, cam_dirs)
                return base, [sort_by_last_number(n) for n in names]
            except Exception:
                continue
        raise ValueError(f"No json files found in {self.pose_dir} nor {self.sync_dir} subdirectories. "
                         f"Make sure you run Pose2Sim.poseEstimation() first.")


class NativeIndex:
    """The camera folders of one input directory listed, filtered, sorted and tabulated by the native code
    (`p2s_index_*`, csrc/p2s_json.cpp) — the statements below (`_list`, `sort_by_last_number`, `frame_file_table`,
    `frame_paths`) without 40 k Python strings in between."""

    def __init__(self, base, cam_dirs):
        import ctypes as C
        from . import _lib
        self.h = None
        self.lib = _lib.load()
        self.n_cams = len(cam_dirs)
        dirs = [os.path.join(base, d).encode() for d in cam_dirs]
        arr = (C.c_char_p * len(dirs))(*dirs)
        h = C.c_void_p()
        if self.lib.p2s_index_open(C.cast(arr, C.c_void_p), len(dirs), C.byref(h)) != 0:
            raise OSError(f"cannot list the camera folders of {base}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.p2s_index_close(self.h)
            self.h = None

    __del__ = close

    def counts(self):
        return [int(self.lib.p2s_index_file_count(self.h, c)) for c in range(self.n_cams)]

    def names(self):
        """Per camera the sorted file names (what `files_for_triangulation` returns)."""
        return [[self.lib.p2s_index_file_name(self.h, c, i).decode() for i in range(n)] for c, n in enumerate(self.counts())]

    def build_table(self, f_range):
        """False when a listed name holds no number (the Python statements raise the reference's IndexError)."""
        f0, f1 = (int(f_range[0]), int(f_range[1])) if len(f_range) >= 2 else (0, int(f_range[0]))
        self.n_frames = len(range(f0, f1))
        return self.lib.p2s_index_build_table(self.h, f0, f1) == 0

    def table_paths(self):
        """The current table as [F][C] Python strings ('' = no file) — for tests; the reader takes the C array."""
        import ctypes as C
        ptr = C.cast(self.lib.p2s_index_table_paths(self.h), C.POINTER(C.c_char_p))
        return [[ptr[f * self.n_cams + c].decode() for c in range(self.n_cams)] for f in range(self.n_frames)]

    def table_names(self):
        """The current table as [F][C] file NAMES ('none' = no file), what `frame_file_table` returns — from ONE copy of
        the native buffer instead of a ctypes call per entry."""
        import ctypes as C
        n = C.c_longlong(0)
        ptr = self.lib.p2s_index_table_arena(self.h, C.cast(C.pointer(n), C.c_void_p))
        flat = C.string_at(ptr, n.value).decode().split("\0")[:self.n_frames * self.n_cams] if n.value else []
        nc = self.n_cams
        return [[(p.rsplit("/", 1)[1] if p else "none") for p in flat[f * nc:(f + 1) * nc]] for f in range(self.n_frames)]

    def table_c_array(self):
        """The table's `const char *const *` for the native readers / writers (valid until the next build_table / close)."""
        return self.lib.p2s_index_table_paths(self.h)

    def signature(self):
        import ctypes as C
        sig = (C.c_ulonglong * 2)()
        if self.lib.p2s_index_signature(self.h, C.cast(sig, C.c_void_p), 0) != 0:
            raise OSError("signature")
        return f"{sig[0]:016x}{sig[1]:016x}"

    def read(self, keypoints_ids, nb_persons, n_threads=0):
        """`read_pose_files` on the current table."""
        import ctypes as C
        from . import _lib
        F, n_cams = self.n_frames, self.n_cams
        K, N = len(keypoints_ids), int(nb_persons)
        ids = np.ascontiguousarray(keypoints_ids, dtype=np.int32)
        x = np.empty((F, N, K, n_cams), np.float32)
        y, lik = np.empty_like(x), np.empty_like(x)
        if F == 0:
            return x, y, lik, 0
        n_people = np.empty((F, n_cams), np.int32)
        inexact = C.c_longlong(0)
        _lib.check(None, self.lib.p2s_read_pose_files(self.lib.p2s_index_table_paths(self.h), F, n_cams, ids.ctypes.data, K, N,
                                                      x.ctypes.data, y.ctypes.data, lik.ctypes.data, n_people.ctypes.data, None,
                                                      C.cast(C.pointer(inexact), C.c_void_p), int(n_threads)))
        return x, y, lik, int(inexact.value)


def frame_file_table(json_files_names, f_range):
    """File name per (frame, camera) with the reference's selection rule (:799-800) in O(F):
    every camera contributes ALL its files whose last number is f, or 'none'; the lists are then
    flattened and the first n_cams entries are used (so a camera with two files for one frame shifts
    the later cameras, exactly as in the reference)."""
    n_cams = len(json_files_names)
    per_cam = []
    for names in json_files_names:
        d = {}
        for j in names:
            d.setdefault(frame_number(j), []).append(j)
        per_cam.append(d)
    table = []
    for f in range(*f_range):
        flat = []
        for c in range(n_cams):
            flat.extend(per_cam[c].get(f) or ["none"])
        table.append(flat[:n_cams])
    return table


def load_json(path):
    """Parsed file or None when it cannot be opened / parsed (the reference's bare `except:`)."""
    try:
        with open(path, "r") as f:
            return json.load(f)
    except Exception:
        return None


def _person_keypoints(js, n):
    """`js['people'][n]['pose_keypoints_2d']` as a float64 vector, or None on any failure."""
    try:
        kp = js["people"][n]["pose_keypoints_2d"]
        return np.asarray(kp, dtype=np.float64).reshape(-1)
    except Exception:
        return None


def gather_keypoints(kp, ids3):
    """x, y, likelihood of the keypoints at flat offsets ids3 (= 3 * id); NaN where out of range."""
    K = len(ids3)
    out = np.full((3, K), np.nan)
    if kp is None:
        return out
    ok = ids3 + 2 < kp.shape[0]
    if ok.all():
        out[0], out[1], out[2] = kp[ids3], kp[ids3 + 1], kp[ids3 + 2]
    else:
        i = ids3[ok]
        out[0, ok], out[1, ok], out[2, ok] = kp[i], kp[i + 1], kp[i + 2]
    return out


def float32_inexact(*arrays):
    """Number of finite values that change when rounded to float32."""
    n = 0
    for a in arrays:
        with np.errstate(invalid="ignore", over="ignore"):
            n += int(np.count_nonzero(np.isfinite(a) & (a.astype(np.float32).astype(np.float64) != a)))
    return n


def frame_paths(input_dir, cam_dirs, table):
    """Absolute path per (frame, camera); '' where the frame has no file ('none' in the reference)."""
    prefix = [os.path.join(input_dir, d, "") for d in cam_dirs]          # one join per camera, not per file
    n = len(cam_dirs)
    return [[prefix[c] + names[c] if names[c] != "none" else "" for c in range(n)] for names in table]


def read_pose_files(paths, keypoints_ids, nb_persons, n_threads=0):
    """Native multi-threaded reader (csrc/p2s_json.cpp, `p2s_read_pose_files`): every file parsed once.
    paths: [F][C] strings.  Returns x, y, lik float32 [F, N, K, C], n_people int32 [F, C], n_inexact."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    F = len(paths)
    n_cams = len(paths[0]) if F else 1
    K, N = len(keypoints_ids), int(nb_persons)
    flat = [p.encode() for row in paths for p in row]
    arr = (C.c_char_p * len(flat))(*flat)
    ids = np.ascontiguousarray(keypoints_ids, dtype=np.int32)
    x = np.empty((F, N, K, n_cams), np.float32)
    y = np.empty_like(x)
    lik = np.empty_like(x)
    n_people = np.empty((F, n_cams), np.int32)
    inexact = C.c_longlong(0)
    _lib.check(None, lib.p2s_read_pose_files(C.cast(arr, C.c_void_p), F, n_cams, ids.ctypes.data, K, N, x.ctypes.data,
                                             y.ctypes.data, lik.ctypes.data, n_people.ctypes.data, None,
                                             C.cast(C.pointer(inexact), C.c_void_p), int(n_threads)))
    return x, y, lik, n_people, int(inexact.value)


# ---- staging cache: parsed trials as memory-mapped SoA planes -------------------------------------------------------
# once the search takes microseconds, parsing F x C small JSON files is the stage.  A parsed trial
# is therefore kept as ONE file of float32 planes [3][F][N][K][C] (the layout the device call takes) under
# $P2S_CACHE_DIR (default ~/.cache/pose2sim_b200) and memory-mapped on the next run.  The key is a digest of every file's
# path, mtime (ns) and size (one native multi-threaded stat pass, `p2s_stat_files`) plus the keypoint ids and the person
# count, so any added, removed, rewritten or touched file misses.  P2S_STAGE_CACHE=0 turns it off; trials below
# CACHE_MIN_FILES are not worth a cache entry; at most CACHE_MAX_ENTRIES entries are kept (oldest access first out).
CACHE_MIN_FILES = 2000
CACHE_MAX_ENTRIES = 32


def cache_dir():
    return os.environ.get("P2S_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "pose2sim_b200")


def stat_files(paths_flat, n_threads=0):
    import ctypes as C
    from . import _lib
    arr = (C.c_char_p * len(paths_flat))(*[p.encode() for p in paths_flat])
    mt, sz = np.empty(len(paths_flat), np.int64), np.empty(len(paths_flat), np.int64)
    _lib.check(None, _lib.load().p2s_stat_files(C.cast(arr, C.c_void_p), len(paths_flat), mt.ctypes.data, sz.ctypes.data, int(n_threads)))
    return mt, sz


def staging_key(paths, keypoints_ids, nb_persons):
    import hashlib
    flat = [p for row in paths for p in row]
    mt, sz = stat_files(flat)
    h = hashlib.sha1(b"p2s-staging-v1")
    h.update("\n".join(flat).encode())
    h.update(mt.tobytes())
    h.update(sz.tobytes())
    h.update(np.asarray(keypoints_ids, np.int64).tobytes())
    h.update(str((int(nb_persons), len(paths), len(paths[0]) if paths else 0)).encode())
    return h.hexdigest()


def _cache_trim(d):
    try:
        entries = sorted((e for e in os.scandir(d) if e.name.endswith(".npy")), key=lambda e: e.stat().st_atime)
        for e in entries[:max(0, len(entries) - CACHE_MAX_ENTRIES)]:
            os.remove(e.path)
    except OSError:
        pass


def _cache_lookup(key):
    d = cache_dir()
    os.makedirs(d, exist_ok=True)
    entry = os.path.join(d, key)
    try:
        planes = np.load(entry + ".npy", mmap_mode="r")
        inexact = int(open(entry + ".txt").read())
        os.utime(entry + ".npy")
        return entry, (planes[0], planes[1], planes[2], inexact)
    except (OSError, ValueError):
        return entry, None


def _cache_store(entry, x, y, lik, inexact):
    try:
        tmp = entry + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            np.save(f, np.stack([x, y, lik]))
        with open(entry + ".txt", "w") as f:
            f.write(str(inexact))
        os.replace(tmp, entry + ".npy")
        _cache_trim(os.path.dirname(entry))
    except OSError:
        pass


def stage_triangulation_indexed(index, f_range, keypoints_ids, nb_persons):
    """`stage_triangulation` on a NativeIndex: table, cache signature and parse all in native code.  Returns None when
    the Python statements must run instead (a listed name without a number: they raise the reference's IndexError)."""
    if not index.build_table(f_range):
        return None
    n_files = index.n_frames * index.n_cams
    entry = None
    if os.environ.get("P2S_STAGE_CACHE", "1") != "0" and n_files >= CACHE_MIN_FILES:
        try:
            import hashlib
            key = hashlib.sha1((index.signature() + str((list(keypoints_ids), int(nb_persons)))).encode()).hexdigest()
            entry, hit = _cache_lookup(key)
            if hit is not None:
                return hit
        except OSError:
            entry = None
    x, y, lik, inexact = index.read(keypoints_ids, nb_persons)
    if entry is not None:
        _cache_store(entry, x, y, lik, inexact)
    return x, y, lik, inexact


def stage_triangulation(input_dir, cam_dirs, json_files_names, f_range, keypoints_ids, nb_persons):
    """All frames of `extract_files_frame_f` (triangulation.py:607-653) at once, through the native reader.
    Returns x, y, lik float32 [F, N, K, C] (unit-major, camera fastest: the layout `p2s_triangulate_host`
    takes viewed as [F*N*K, C]) and the number of values float32 could not represent exactly.  Large trials come
    from / go to the staging cache (above)."""
    table = frame_file_table(json_files_names, f_range)
    paths = frame_paths(input_dir, cam_dirs, table)
    n_files = len(paths) * (len(paths[0]) if paths else 0)
    use_cache = os.environ.get("P2S_STAGE_CACHE", "1") != "0" and n_files >= CACHE_MIN_FILES
    entry = None
    if use_cache:
        try:
            d = cache_dir()
            os.makedirs(d, exist_ok=True)
            entry = os.path.join(d, staging_key(paths, keypoints_ids, nb_persons))
            planes = np.load(entry + ".npy", mmap_mode="r")
            inexact = int(open(entry + ".txt").read())
            os.utime(entry + ".npy")
            return planes[0], planes[1], planes[2], inexact
        except (OSError, ValueError):
            pass
    x, y, lik, _, inexact = read_pose_files(paths, keypoints_ids, nb_persons)
    if entry is not None:
        try:
            tmp = entry + f".tmp{os.getpid()}"
            with open(tmp, "wb") as f:
                np.save(f, np.stack([x, y, lik]))
            with open(entry + ".txt", "w") as f:
                f.write(str(inexact))
            os.replace(tmp, entry + ".npy")
            _cache_trim(os.path.dirname(entry))
        except OSError:
            pass
    return x, y, lik, inexact


def stage_triangulation_python(input_dir, cam_dirs, json_files_names, f_range, keypoints_ids, nb_persons):
    """Pure-Python twin of `stage_triangulation` (json.load per file), kept as the restatement the native
    parser is tested against (tests/test_native_staging.py).  Returns float64 arrays [F, N, K, C]."""
    table = frame_file_table(json_files_names, f_range)
    F, C, K, N = len(table), len(cam_dirs), len(keypoints_ids), nb_persons
    ids3 = 3 * np.asarray(keypoints_ids, dtype=np.int64)
    x = np.full((F, N, K, C), np.nan)
    y = np.full((F, N, K, C), np.nan)
    lik = np.full((F, N, K, C), np.nan)
    for fi, names in enumerate(table):
        for c in range(C):
            js = load_json(os.path.join(input_dir, cam_dirs[c], names[c]))
            if js is None:
                continue
            for n in range(N):
                v = gather_keypoints(_person_keypoints(js, n), ids3)
                x[fi, n, :, c], y[fi, n, :, c], lik[fi, n, :, c] = v[0], v[1], v[2]
    return x, y, lik


def count_persons(input_dir, cam_dirs, json_files_names):
    """triangulation.py:784 + :77-90: the largest `len(people)` over EVERY listed json of every camera
    (native reader, one parse per file).  A file that cannot be parsed raises, as in the reference."""
    best = 0
    for c, names in enumerate(json_files_names):
        paths = [[os.path.join(input_dir, cam_dirs[c], name)] for name in names]
        if not paths:
            continue
        _, _, _, n_people, _ = read_pose_files(paths, [], 0)
        if (n_people < 0).any():
            bad = os.path.join(input_dir, cam_dirs[c], names[int(np.flatnonzero(n_people[:, 0] < 0)[0])])
            with open(bad) as f:                      # the reference's own statement (:88-89): the same FileNotFoundError /
                json.load(f)                          # json.JSONDecodeError (a ValueError) with the same position comes out
            raise ValueError(f"cannot parse {bad}")
        best = max(best, int(n_people.max(initial=0)))
    return best


# ---- association: native reader / writer ---------------------------------------------------------------------
def _c_paths(paths):
    import ctypes as C
    flat = [p.encode() for row in paths for p in row]
    return (C.c_char_p * len(flat))(*flat)


def read_people_files(paths, value_offset, n_values, max_persons, n_threads=0, c_array=None):
    """Native reader of the association stage (`p2s_read_people_files`): paths [F][C] -> obs float32
    [F, C, max_persons, n_values] (None when n_values == 0), count_named, count_listed, list_len int32 [F, C],
    status uint8 [F, C] and the number of values float32 could not hold exactly."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    if isinstance(paths, tuple):                              # (n_frames, n_cams): the paths are in `c_array`
        F, n_cams = paths
    else:
        F = len(paths)
        n_cams = len(paths[0]) if F else 1
    arr = c_array if c_array is not None else _c_paths(paths)
    obs = np.empty((F, n_cams, max_persons, n_values), np.float32) if n_values * max_persons else None
    named, listed, llen = (np.zeros((F, n_cams), np.int32) for _ in range(3))
    status = np.zeros((F, n_cams), np.uint8)
    inexact = C.c_longlong(0)
    _lib.check(None, lib.p2s_read_people_files(arr if isinstance(arr, int) else C.cast(arr, C.c_void_p), F, n_cams, int(value_offset), int(n_values), int(max_persons),
                                               obs.ctypes.data if obs is not None else None, named.ctypes.data, listed.ctypes.data,
                                               llen.ctypes.data, status.ctypes.data, C.cast(C.pointer(inexact), C.c_void_p), int(n_threads)))
    return obs, named, listed, llen, status, int(inexact.value)


def rewrite_people_files(src_paths, dst_paths, proposals, n_threads=0):
    """Native writer of the association stage (`p2s_rewrite_people_files`).  proposals[f] = array [n_f, C] of person
    indices (NaN = camera off).  Returns status uint8 [F, C] (1 written, 0 no file, 2 needs the Python writer)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    if isinstance(src_paths, list):
        F = len(src_paths)
        n_cams = len(src_paths[0]) if F else 1
    else:
        F, n_cams = len(proposals), len(dst_paths[0]) if len(dst_paths) else 1
    if isinstance(proposals, np.ndarray) and proposals.ndim == 3 and proposals.shape[0] == F:
        # the same number of proposals in every frame (single-person mode: one): no per-frame loop
        a = np.asarray(proposals, dtype=np.float64).reshape(F, -1, n_cams)
        comb = np.ascontiguousarray(np.where(np.isnan(a), -1, a).astype(np.int32).reshape(-1, n_cams))
        offs = (np.arange(F + 1, dtype=np.int64) * a.shape[1]).astype(np.int32)
    else:
        offs = np.zeros(F + 1, np.int32)
        rows = []
        for f, prop in enumerate(proposals):
            a = np.asarray(prop, dtype=np.float64).reshape(-1, n_cams) if np.size(prop) else np.zeros((0, n_cams))
            rows.append(np.where(np.isnan(a), -1, a).astype(np.int32))
            offs[f + 1] = offs[f] + len(a)
        comb = np.ascontiguousarray(np.concatenate(rows) if rows else np.zeros((0, n_cams), np.int32), dtype=np.int32)
    status = np.zeros((F, n_cams), np.uint8)
    s_arr = src_paths if not isinstance(src_paths, list) else _c_paths(src_paths)      # a ctypes array is taken as it is
    d_arr = dst_paths if not isinstance(dst_paths, list) else _c_paths(dst_paths)
    _lib.check(None, lib.p2s_rewrite_people_files(s_arr if isinstance(s_arr, int) else C.cast(s_arr, C.c_void_p),
                                                  d_arr if isinstance(d_arr, int) else C.cast(d_arr, C.c_void_p), F, n_cams, offs.ctypes.data,
                                                  comb.ctypes.data if len(comb) else None, status.ctypes.data, int(n_threads)))
    return status


# ---- association --------------------------------------------------------------------------------------
def persons_per_camera(js):
    """personAssociation.py:81-89: people whose x values are not all NaN; any failure -> 0."""
    try:
        people = js["people"]
        return len([p for p in people if not all(np.isnan(p["pose_keypoints_2d"][::3]))])
    except Exception:
        return 0


def read_people(js):
    """personAssociation.py:260-274 `read_json`: keypoint lists with at least 3 values; any failure -> []."""
    try:
        out = []
        for p in js["people"]:
            if len(p["pose_keypoints_2d"]) < 3:
                continue
            out.append(p["pose_keypoints_2d"])
        return out
    except Exception:
        return []


def stage_association(read_dir, cam_dirs, table, tracked_keypoint_id, max_persons):
    """Tracked-keypoint observations of every detected person: obs [F, C, max_persons, 4] float32
    {x, y, likelihood, 0} and count [F, C] int32.

    Two index spaces of the reference are kept apart: `count` comes from `persons_combinations`
    (people with a non-NaN x), while the observation of person index p is taken from `read_json`'s
    list (people with >= 3 values); an index beyond that list is a NaN triple (the `except` at
    personAssociation.py:203-205).  Raises ValueError when a camera shows more than `max_persons`."""
    F, C = len(table), len(cam_dirs)
    obs = np.full((F, C, max_persons, 4), np.nan, np.float64)
    obs[..., 3] = 0.0
    count = np.zeros((F, C), np.int32)
    parsed = []
    t3 = 3 * int(tracked_keypoint_id)
    for fi, names in enumerate(table):
        row = []
        for c in range(C):
            js = load_json(os.path.join(read_dir, cam_dirs[c], names[c]))
            row.append(js)
            if js is None:
                continue
            n = persons_per_camera(js)
            if n > max_persons:
                raise ValueError(f"{n} persons in {names[c]}: the device search handles at most {max_persons} per camera")
            count[fi, c] = n
            people = read_people(js)
            for p in range(min(n, len(people))):
                try:
                    v = [float(t) for t in people[p][t3:t3 + 3]]
                except Exception:
                    continue
                if len(v) == 3:
                    obs[fi, c, p, :3] = v
        parsed.append(row)
    return obs, count, parsed
