"""CPU oracle for the EEG-distillation hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / torch-CPU restatement of the reference
algorithms on the hot path (band-pass filter, LSTM encoder, DINO head, DINO
loss + centre EMA, Adam step).  It exists to CHECK the CUDA product path in
`cerebralsignalnetworks_b200/`; the product never imports it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import anything from here.

Pinning status (see DESIGN.md "Oracle"):
  * DINOLoss (single-view + multi-crop), DINOHead, MultiCropWrapper,
    cosine_scheduler, get_params_groups, clip_gradients: pinned against the
    reference's OWN classes, executed in the build container through
    `oracle/ref_import.py`; golden vectors committed under `tests/golden/`
    (generator: `oracle/make_golden.py`).
  * Filter: pinned against scipy.signal (the arithmetic the reference calls at
    utils/Utilities.py:421-427 and designs at utils/EEGFilters.py:26-39).
  * LSTM encoder `models.lstm.Model`: the reference does NOT ship this file
     -> restated from the call sites on
    torch.nn.LSTM; *parity unpinned by the reference* for this one class.
"""
