"""Callbacks: host-side mirror of pyspeedy/callbacks.py (same classes, arguments and behaviour)."""
import copy
import os

from pyspeedy_b200 import DEFAULT_OUTPUT_VARS
from pyspeedy_b200.dataset import Dataset
from pyspeedy_b200.speedy import Speedy, SpeedyEns  # noqa


class BaseCallback:
    """Base callback class (pyspeedy/callbacks.py:31-75)."""

    def __init__(self, *args, **kwargs):
        self.verbose = kwargs.pop("verbose", False)
        self.interval = kwargs.pop("interval", 1)
        self.spinup_date = kwargs.pop("spinup_date", None)

    def skip_flag(self, model_instance):
        if self.spinup_date is not None:
            if model_instance.current_date < self.spinup_date:
                return True
        return model_instance.get_current_step() % self.interval != 0

    def print_msg(self, msg):
        if self.verbose:
            print(msg)

    def copy(self):
        return copy.deepcopy(self)

    def __call__(self, model_instance):
        pass


class DiagnosticCheck(BaseCallback):
    """Check that the prognostic variables are inside reasonable ranges (pyspeedy/callbacks.py:78-112)."""

    def __init__(self, interval=36):
        super().__init__(interval=interval)

    def __call__(self, model_instance):
        if self.skip_flag(model_instance):
            return
        # a failing check raises a RuntimeError; for an ensemble the members are checked by ONE batched driver call
        # (the reference loops over the members: pyspeedy/callbacks.py:103-111)
        model_instance.check()


class ModelCheckpoint(BaseCallback):
    """Keep a time series of selected grid variables in memory (pyspeedy/callbacks.py:115-180)."""

    def __init__(self, interval=36, verbose=False, spinup_date=None, variables=None, output_dir="./"):
        if variables is None:
            variables = DEFAULT_OUTPUT_VARS
        self.variables = variables
        self.output_dir = output_dir
        self.history_interval = interval
        super().__init__(verbose=verbose, interval=interval, spinup_date=spinup_date)
        self._frames, self._merged = [], None

    @property
    def dataframe(self):
        """The time series so far.  The reference re-merges the whole series at every checkpoint (xr.merge of the
        growing dataset, O(n^2) copies); here the checkpoints are kept as they come and merged once, when read."""
        if self._merged is None and self._frames:
            self._merged = self._frames[0] if len(self._frames) == 1 else Dataset.merge(self._frames)
            self._frames = [self._merged]
        return self._merged

    def __call__(self, model_instance):
        if self.skip_flag(model_instance):
            return
        self._frames.append(model_instance.to_dataframe(variables=self.variables))
        self._merged = None


class EnsembleStatistics(BaseCallback):
    """Extension (no reference counterpart; the notebook examples/Ensemble_forecast.ipynb computes these with xarray from a
    ModelCheckpoint of all members): keep the time series of the ensemble mean and spread (std, ddof=0) of the six
    default outputs.  The sums are formed on the GPU in the epilogue of the batched spectral2grid and reduced over all
    ranks of a sharded ensemble by one NCCL all-reduce per output time (``SpeedyEns.mean_and_spread``)."""

    def __init__(self, interval=36, verbose=False, spinup_date=None, variables=None):
        self.variables = DEFAULT_OUTPUT_VARS if variables is None else variables
        super().__init__(verbose=verbose, interval=interval, spinup_date=spinup_date)
        self.times, self.mean, self.spread = [], {v: [] for v in self.variables}, {v: [] for v in self.variables}

    def __call__(self, model_instance):
        if self.skip_flag(model_instance):
            return
        stats = model_instance.mean_and_spread(self.variables)
        self.times.append(model_instance.current_date)
        for v in self.variables:
            self.mean[v].append(stats[v][0])
            self.spread[v].append(stats[v][1])
        self.print_msg(f"Ensemble statistics at {model_instance.current_date}.")


class XarrayExporter(BaseCallback):
    """Write one NetCDF file per output time (pyspeedy/callbacks.py:183-255)."""

    def __init__(self, interval=36, verbose=False, spinup_date=None, variables=None, output_dir="./",
                 filename_fmt="%Y-%m-%d_%H%M.nc"):
        if variables is None:
            variables = DEFAULT_OUTPUT_VARS
        self.variables = variables
        self.output_dir = output_dir
        self.filename_fmt = filename_fmt
        self.history_interval = interval
        super().__init__(verbose=verbose, interval=interval, spinup_date=spinup_date)

    def __call__(self, model_instance):
        if self.skip_flag(model_instance):
            return
        model_df = model_instance.to_dataframe(variables=self.variables)
        file_name = model_instance.current_date.strftime(self.filename_fmt)
        os.makedirs(self.output_dir, exist_ok=True)
        output_file_path = os.path.join(self.output_dir, file_name)
        self.print_msg(f"Saving model output at: {output_file_path}.")
        model_df.to_netcdf(output_file_path, encoding=dict())
