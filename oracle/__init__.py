"""CPU oracle for the BALF inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``balf_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker / the CPU arm.

Every function restates a function of the reference (ericzzj1989/BALF, mounted
read-only at /root/reference while the repo is developed) and cites the
``file:line`` it follows.  Parity status:

* detector, head, pixel-shuffle, HardNet, pad/unpad, border mask, windowed NMS,
  k-th-value top-k, greedy ``nms_fast``, thresholding: PINNED -- checked against
  the reference's own Python modules imported from /root/reference
  (``oracle/make_golden.py`` wrote ``tests/golden/*.npz``; ``tests/test_oracle_*``
  re-check the restatement against those vectors on every run).
* round 2 (``python -m oracle.make_golden r2`` -> ``tests/golden/r2_*.npz``), all PINNED:
  the reference's ``detect()`` on its own fixtures media/im1.jpg / im2.jpg and PIL's
  ``convert('L')`` planes; reference score maps at 960x1216 and 1024x1024 and ``detect()`` at
  900x1200; reference HardNet on 2048 patches; the repeatability metrics and homography helpers
  of balf/benchmark_test (``oracle/metrics.py``; ``cv2.warpPerspective`` restated and checked
  against OpenCV itself); ``box_nms`` against ``torchvision.ops.nms``; the seeded weights of
  ``oracle/weights.py`` against the reference constructors' digests.
* sub-pixel soft-argmax (torchgeometry 0.1.2), LAF / patch pyramid sampling and
  SMNN matching (kornia 0.7.4): PARITY UNPINNED -- the two pip dependencies are
  not vendored in the reference and are not installable offline; their published
  algorithms are restated in ``oracle/thirdparty.py`` and anchored only on the
  reference's call sites.
"""
