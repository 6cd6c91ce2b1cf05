"""f4 on the GPU (sygnals/core/ml_utils/scaling.py:49-175): fit_scaler / apply_scaling on CUDA tensors -- the feature
vectors where the engine leaves them -- against scikit-learn, incl. the robust scaler, NaN handling and a constant column.
(The two-rank NCCL leg lives in tests/test_dist_nccl.py.)"""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("kind,params", [("standard", {}), ("standard", {"with_mean": False}), ("minmax", {"feature_range": (-1.0, 2.0)}),
                                         ("robust", {}), ("robust", {"quantile_range": (5.0, 95.0), "with_centering": False})])
def test_scalers_on_cuda_vs_sklearn(kind, params):
    import torch
    from sklearn.preprocessing import MinMaxScaler, RobustScaler, StandardScaler
    from sygnals_b200.core.ml_utils import scaling
    rng = np.random.default_rng(3)
    X = rng.standard_normal((5000, 24)) * 10.0 ** rng.integers(-3, 3, 24) + rng.standard_normal(24)
    X[:, 7] = 2.5
    X[rng.integers(0, 5000, 200), 3] = np.nan
    sk = {"standard": StandardScaler, "minmax": MinMaxScaler, "robust": RobustScaler}[kind](**params).fit(X)
    Xd = torch.from_numpy(X).cuda()
    Y, sc = scaling.apply_scaling(Xd, kind, params)
    assert Y.is_cuda and Y.dtype == torch.float64
    np.testing.assert_allclose(Y.cpu().numpy(), sk.transform(X), rtol=1e-9, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(sc.scale_, sk.scale_, rtol=1e-10)
    # transform-only with the fitted object (apply_scaling(fit=False), scaling.py:121-131)
    Y2, _ = scaling.apply_scaling(Xd[:10], kind, fit=False, scaler_instance=sc)
    np.testing.assert_allclose(Y2.cpu().numpy(), sk.transform(X[:10]), rtol=1e-9, atol=1e-9, equal_nan=True)
    with pytest.raises(ValueError):
        scaling.apply_scaling(Xd, "nope")
