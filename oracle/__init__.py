"""CPU oracle for the flow hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The product
(``ir2rgb_b200``) never imports it and has no CPU fallback.

Two oracles live here:

* :mod:`oracle.c_oracle`   -- ctypes binding of ``flowops_oracle.c``, a literal transliteration of the
  reference CUDA kernels (bit-level op order), numpy in / numpy out.
* :mod:`oracle.torch_ref`  -- the pure-PyTorch restatement the north star names (displacement-loop
  correlation, ``F.grid_sample`` warps, ``sqrt(sum(x*x))``), differentiable, fp32 or fp64.

``oracle/_ref/`` (git-ignored) holds the reference's own CUDA extensions rebuilt for sm_100 by
``oracle/build_ref.py``; :mod:`oracle.ref_ext` loads them when present (GPU only).
"""
