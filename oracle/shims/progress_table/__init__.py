"""Test-infrastructure shim for the third-party `progress_table` package (absent from this image).

Only the surface the reference touches (stage.py:147,159,168,192,195-205): a table that remembers columns and rows.
Used ONLY by oracle/gen_golden.py to run the unmodified reference inside the build container.
"""


class ProgressTable:
    def __init__(self, *args, file=None, **kwargs):
        self.file = file
        self.columns = {}
        self.rows = []
        self._row = {}

    def add_column(self, name, **kwargs):
        self.columns[name] = dict(kwargs)

    def __setitem__(self, name, value):
        self._row[name] = value

    def __getitem__(self, name):
        return self._row.get(name)

    def update(self, name, value, **kwargs):
        self._row[name] = value

    def next_row(self, **kwargs):
        self.rows.append(self._row)
        self._row = {}

    def close(self):
        if self._row:
            self.next_row()
