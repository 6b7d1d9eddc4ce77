"""CPU oracle for the DeepWMH 3d_fullres sliding-window inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepwmh_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline.

PARITY UNPINNED: the arithmetic of this path lives in the un-vendored
``nnunet`` package of ``lchdl/nnUNet_for_DeepWMH`` (fork of MIC-DKFZ/nnUNet v1,
no pinned version).  ``/root/reference`` holds no
golden vectors, known-answer tests or fixtures for it, and the package cannot
be imported in the authoring container.  The oracle is therefore a restatement
of the published nnU-Net v1 algorithm from stock ``torch.nn`` ops, anchored on
the reference's call sites (``deepwmh/main/predict.py:133-156``,
``deepwmh/pipeline/DCNN_multistage.py:331-344,531-535``) and checked against
the known-answer properties of SURVEY.md section 8c (``tests/test_oracle.py``).
"""
from .nnunet_oracle import *  # noqa: F401,F403
