"""
oracle/ -- TEST INFRASTRUCTURE ONLY.  Not shipped, not a fallback.

CPU (numpy/scipy, float64) restatement of the Sygnals segment->features hot
path, used solely as the *checker* for the CUDA engine in ``sygnals_b200``:

* ``oracle.librosa_shim``   restates the ~14 librosa (PyPI ``librosa``,
  reference pin ``>=0.10.0``, developed against 0.11.0 --
  ``/root/reference/pyproject.toml:38``) entry points the reference calls on
  this path.  librosa is NOT vendored under ``/root/reference`` and is NOT
  installable here (no network), so its published algorithm is restated.
* ``oracle.sygnals_oracle``  restates the reference's own functions
  (``sygnals/core/{dsp,segmentation}.py``, ``sygnals/core/features/*.py``,
  ``sygnals/core/audio/features.py``) on top of the shim, each function citing
  the reference file:line it follows.  It travels to the GPU box.
* ``oracle.ref_loader``      (this container only) imports the UNMODIFIED
  reference from ``/root/reference`` with ``librosa`` replaced by the shim.
  It generates ``tests/golden/*.npz`` (``tests/golden/make_golden.py``) and
  validates ``sygnals_oracle`` line by line.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``sygnals_b200`` never
does; its product path raises if the CUDA library is missing.

PARITY PIN STATUS
-----------------
* sygnals' own code (framing arithmetic, per-frame centroid/rolloff/crest,
  segmentation, naming, length fix-ups): PINNED -- the unmodified reference
  runs here (``ref_loader``) and its outputs are committed as golden vectors;
  the reference's own known-answer tests (tests/test_segmentation.py,
  tests/test_features_time.py, tests/test_features_freq.py) pass against it
  and are restated in ``tests/test_oracle_known_answers.py``.
* the librosa layer (STFT/mel/power_to_db/DCT/spectral_contrast/rms):
  "parity unpinned" against a real librosa install (none available offline and
  the reference's tests hold shapes/dtypes only, no numeric golden vectors for
  it).  The shim is cross-checked against independent in-image
  implementations instead: ``torch.stft`` (<=1e-12), ``torchaudio`` slaney mel
  filterbank (<=2e-7, float32 rounding), ``scipy.fftpack.dct`` (the very call
  librosa makes), ``scipy.signal.get_window``.
"""
