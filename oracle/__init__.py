"""CPU oracle for the LWSNet stereo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``lwsnet_b200/`` (the product) may import
this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / the reported CPU baseline.

PARITY UNPINNED: the reference (PrinceVictor/LWSNet) is pure Python on
PaddlePaddle 2.0.0rc0 (paddle_env.yml:149).  Paddle is not installable in this
image (no wheel, no network), the reference ships no tests, golden vectors or
known-answer fixtures, and its only sample outputs
(reference/1..4.png) are uint8/JET renders of a Google-Drive checkpoint.  The
oracle therefore restates models/models.py and models/submodules.py line by
line on PyTorch-CPU + NumPy with the Paddle operator semantics listed in
SURVEY.md Appendix C, and is cross-checked only against (a) an independent
closed-form NumPy specification of each hot-path function (oracle/spec_np.py)
and (b) torch's own grid_sample / interpolate kernels.
"""
