"""Egress: the files written from engine results load back as the reference's read_data / save dataset expect
(sygnals/core/data_handler.py:160-169, 245-275; sygnals/cli/save_cmd.py:140-190)."""
import numpy as np
import pandas as pd
import pytest

from sygnals_b200.core import data_handler as dh


def test_feature_rows_npz_and_csv(tmp_path):
    names = ["mfcc_0", "mfcc_1", "rms_energy"]
    rows = np.arange(12, dtype=np.float32).reshape(3, 4)
    times = np.arange(4) * 0.5
    dh.save_feature_rows(tmp_path / "f.npz", names, rows, times)
    with np.load(tmp_path / "f.npz") as z:                      # what read_data does (data_handler.py:163-164)
        d = {k: z[k] for k in z.files}
    assert list(d) == ["time"] + names and all(v.ndim == 1 and v.dtype == np.float64 for v in d.values())
    np.testing.assert_array_equal(d["mfcc_1"], rows[1])
    # 'vectors' assembly input: a dict of 1-D float arrays (save_cmd.py:163-166) -> usable as is
    assert {k: v for k, v in d.items() if v.ndim == 1}.keys() == d.keys()
    dh.save_feature_rows(tmp_path / "f.csv", names, rows, times)
    df = pd.read_csv(tmp_path / "f.csv")
    assert list(df.columns) == ["time"] + names and len(df) == 4
    with pytest.raises(ValueError):
        dh.save_feature_rows(tmp_path / "f.bin", names, rows)
    with pytest.raises(ValueError):
        dh.save_feature_rows(tmp_path / "g.npz", names[:2], rows)


def test_segment_vectors_npz_and_csv(tmp_path):
    names = ["a", "b"]
    v = np.array([[1.0, 2.0], [3.0, np.nan], [5.0, 6.0]])
    dh.save_segment_vectors(tmp_path / "d.npz", names, v, labels=["x", "y", "z"])
    with np.load(tmp_path / "d.npz") as z:
        assert z["data"].shape == (3, 2) and z["data"].dtype == np.float64          # save_data(ndarray) -> key 'data' (data_handler.py:257-259)
        np.testing.assert_array_equal(z["data"], v)
        assert list(z["feature_names"]) == names and list(z["labels"]) == ["x", "y", "z"]
    dh.save_segment_vectors(tmp_path / "d.csv", names, v)
    df = pd.read_csv(tmp_path / "d.csv")
    assert list(df.columns) == names and df.shape == (3, 2) and np.isnan(df["b"][1])


@pytest.mark.gpu
def test_writers_take_device_tensors(tmp_path):
    import torch
    t = torch.arange(6, dtype=torch.float32, device="cuda").reshape(2, 3)
    dh.save_feature_rows(tmp_path / "g.npz", ["p", "q"], t)
    with np.load(tmp_path / "g.npz") as z:
        np.testing.assert_array_equal(z["q"], [3.0, 4.0, 5.0])
    dh.save_segment_vectors(tmp_path / "h.npz", ["p", "q", "r"], t.double())
    with np.load(tmp_path / "h.npz") as z:
        assert z["data"].shape == (2, 3)
