"""CPU oracle for the LongTerm360FoV sequence-prediction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``longterm360fov_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or as
the timed CPU baseline.

Parity status
-------------
* Featuriser / windowing / sample builders / evaluation metric (``get_gt_target_xyz``,
  ``get_gt_target_xyz_oth``, ``reshape2second_stacks`` incl. ``purelly_testing``,
  ``generate_fake_batch_numpy``, ``get_whole_span``, ``xyz2thetaphi`` + the one-hot
  index / ``_create_one_hot`` path, the FoV hit rate) are PINNED: the golden vectors in
  ``tests/golden/reference_numpy_golden.npz`` were produced by executing the
  reference's own function bodies from ``/root/reference/mycode/{utility,dataIO,
  others_LSTM_span_whole,baseline_knn_mean}.py`` (see
  ``tests/golden/make_reference_golden.py``).  Likewise ``get_data`` (``reference_get_data_golden.npz``), the
  Gaussian-FoV / head-direction tiles (``reference_gaussian_fov_golden.npz``) and the host utilities
  ``rand_sample_ind`` / ``rand_sample`` / ``clip_xyz`` (``reference_host_utils_golden.npz``), each with its
  ``make_*_golden.py``.
* The layer numerics (LSTM, ConvLSTM2D, Dense, Conv, losses, optimisers) live in
  un-vendored, un-pinned Keras 2.2.x / TensorFlow 1.x which cannot be installed
  here, and the reference ships no tests or golden vectors for them:
  **parity unpinned** for those.  They are anchored instead by (i) two independent
  restatements (``keras_numpy`` loops/einsum vs ``keras_torch`` F.conv2d/autograd)
  agreeing to 1e-6, (ii) a cross-check of the gate algebra/weight layout against
  ``torch.nn.LSTM`` in ``recurrent_activation='sigmoid'`` mode, (iii)
  hand-computable known-answer cases, (iv) float64 finite-difference gradient
  checks, (v) the convolutions and one ConvLSTM2D step against
  ``scipy.signal.correlate2d`` on explicitly padded planes (even and dilated kernels).
* The Philox stream is pinned to the Random123 known-answer vectors; the HDF5 reader of
  ``longterm360fov_b200/h5lite.py`` (product code, not part of this package) to a file libhdf5 wrote.
"""
