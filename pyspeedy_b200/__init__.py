"""pyspeedy_b200 -- B200-native drop-in for the pySPEEDY hot path.

Same public surface as the reference package (pyspeedy/__init__.py): ``Speedy``, ``SpeedyEns``,
``example_bc_file``, ``example_sst_anomaly_file``, ``DEFAULT_OUTPUT_VARS`` and the ``_speedy`` driver object,
which here is a ctypes binding of libspeedy_b200.so (CUDA, sm_100a) instead of the f2py Fortran module.
"""
from pathlib import Path

from ._driver import speedy_driver as _speedy  # noqa

PACKAGE_DATA_DIR = Path(__file__).parent / "data"

DEFAULT_OUTPUT_VARS = ("u_grid", "v_grid", "t_grid", "q_grid", "phi_grid", "ps_grid")


def example_bc_file():
    """Path of the example boundary-conditions file (converted from the reference's example_bc.nc)."""
    return str(PACKAGE_DATA_DIR / "example_bc.npz")


def example_sst_anomaly_file():
    """Path of the example SST anomaly file.  The reference's sst_anomaly.nc is not redistributed here; when the
    file is absent ``Speedy.set_bc`` uses a zero anomaly (documented deviation)."""
    return str(PACKAGE_DATA_DIR / "sst_anomaly.npz")


from .speedy import Speedy, SpeedyEns, MODEL_STATE_DEF  # noqa


def set_sppt(on, seed=0):
    """Switch SPPT (stochastically perturbed parametrisation tendencies, speedy.f90/sppt.f90) on or off for every model
    instance of this process -- the run-time form of the reference's compile-time ``sppt_on`` (params.f90:44)."""
    return _speedy.set_sppt(on, seed)
