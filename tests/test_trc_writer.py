"""The native TRC body writer against pandas' `DataFrame.to_csv` (what the reference calls): byte-identical
text for the same values, including NaN (empty field), integers-as-floats, tiny / huge magnitudes."""
import ctypes as C
import io

import numpy as np
import pandas as pd

from pose2sim_b200 import _lib


def _native(tmp_path, frames, t, vals):
    path = str(tmp_path / "body.trc")
    open(path, "w").close()
    fr = np.ascontiguousarray(frames, np.int64)
    tt = np.ascontiguousarray(t, np.float64)
    v = np.ascontiguousarray(vals, np.float64)
    _lib.check(None, _lib.load().p2s_write_trc_rows(path.encode(), fr.ctypes.data, tt.ctypes.data, v.ctypes.data, len(fr), v.shape[1]))
    return open(path).read()


def _pandas(frames, t, vals):
    df = pd.DataFrame(vals, index=pd.Index(frames))
    df.insert(0, "t", t)
    buf = io.StringIO()
    df.to_csv(buf, sep="\t", index=True, header=None, lineterminator="\n")
    return buf.getvalue()


def test_trc_rows_byte_identical_to_pandas(tmp_path):
    rng = np.random.default_rng(0)
    n, cols = 400, 78
    vals = rng.normal(0, 2, (n, cols))
    vals[rng.random((n, cols)) < 0.05] = np.nan
    vals[0, :12] = [0.0, -0.0, 1.0, -2.0, 1e-5, 1.5e-7, 123456789.0, 1e15, 1e16, 1.2345678901234567e22, 5e-324, -1e-4]
    vals[1, :6] = [0.1, 0.2 + 0.1, 1 / 3, 2 ** 53, 2 ** 53 + 2, 9.999999999999999e15]
    vals[2] = np.float32(rng.normal(0, 1, cols)).astype(np.float64)       # float32-valued doubles
    frames = np.arange(17, 17 + n)
    t = frames / 60
    assert _native(tmp_path, frames, t, vals) == _pandas(frames, t, vals)
    t30 = frames / 30
    assert _native(tmp_path, frames, t30, vals[:, :3]) == _pandas(frames, t30, vals[:, :3])
