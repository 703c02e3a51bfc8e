"""Error codes of the driver (speedy.f90/error_codes.f90:7-9; messages as in pyspeedy/error_codes.py)."""
from collections import defaultdict

ERROR_CODES = defaultdict(lambda: "Unexpected error.")
ERROR_CODES[0] = "Run successful."
ERROR_CODES[-1] = "The model state was not initialized. Please initialize it before running the model."
ERROR_CODES[-2] = "Model variables out of accepted range."
