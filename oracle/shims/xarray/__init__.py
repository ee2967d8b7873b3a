"""Test-infrastructure shim for `xarray` (absent): the reference only names these types in annotations
(util/data.py:71,84,153). Used ONLY by oracle/gen_golden.py."""


class Dataset:
    pass


class DataArray:
    pass
