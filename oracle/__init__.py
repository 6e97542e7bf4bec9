"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product (``hifihr_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Pinning status:
  * MANO (``oracle.mano``): pinned — checked against the reference's own
    ``ManoLayer`` (utils/my_mano.py:225-483) imported unmodified in the build
    container (``oracle/ref_mano.py``) and against SURVEY.md Appendix C known
    answers; golden vectors in ``tests/golden/mano_*.npz``.
  * Rasterizer / shaders / blending / textures (``oracle.p3d``,
    ``oracle/raster_naive.c``): PARITY UNPINNED — PyTorch3D (un-pinned git HEAD,
    README.md:70-71) is a third-party dependency absent from /root/reference and
    from this image; the restatement follows its published algorithm (SURVEY.md
    Appendix A) and is cross-checked only by analytic known-answer tests and by
    two independent restatements (vectorised torch vs scalar C) agreeing.
  * Losses (``oracle.losses``): pinned for SSIM (utils/pytorch_ssim imported
    unmodified in the build container); L1 / mean-RGB / IoU are line-by-line
    restatements of losses.py:355-408 and utils/losses_util.py:366-378.
"""
