"""Drop-in for the reference's vendored `gammatone` package (only `filters` is vendored
there: /root/reference/gammatone/filters.py)."""
from . import filters  # noqa: F401
