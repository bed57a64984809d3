"""CPU oracle for the registration-to-strain hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under the product package may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the reported CPU baseline.

PARITY UNPINNED.  The reference tree (``/root/reference``) does not contain the
arithmetic of this path: the ``models`` package is missing (``main.py:42``) and
``lagomorph`` is an un-vendored, un-pinned third-party dependency
(``README.md:15-17``) that is neither installed nor downloadable here.  The
reference has no tests, golden vectors or fixtures.  This
oracle therefore restates the published lagomorph/PyCA algorithm (SURVEY.md
Appendix A) and is pinned only by (i) the invariants of SURVEY.md section 8c and
(ii) the in-tree pieces that *are* importable: ``modules.loss`` (loss boundary),
``split_vol_to_registration_pairs`` and ``align_n_frames_to`` - see
``tests/golden/make_golden.py``.

Two independent restatements live here and cross-check each other: this torch package (differentiable, any
dtype) and ``lddmm_c.c`` (plain C, own FFT, OpenMP; wrapped by ``oracle.c_oracle``), which is also the
multi-threaded CPU baseline of ``bench.py``.
"""
from .lddmm import (  # noqa: F401
    Conventions,
    DEFAULT,
    interp,
    splat,
    jacobian,
    jacobian_times_vectorfield,
    Ad_star,
    compose_disp_vel,
    FluidMetric,
    EPDiff_step,
    expmap,
    shoot,
)
from .strain import (  # noqa: F401
    N_SECTORS,
    sector_boundaries,
    mask_moments,
    sector_map,
    strain_ecc,
    strain_matrix,
    align_frames,
)
from .path import (  # noqa: F401
    split_vol_to_registration_pairs,
    align_n_frames_to,
    forward_volume,
    forward_pairs,
    registration_reconstruction_loss,
)
from . import augment, path  # noqa: F401,E402
