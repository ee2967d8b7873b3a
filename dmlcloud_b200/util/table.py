"""Per-epoch progress table.  The reference prints through the third-party `progress_table` package
(stage.py:147,159,168,192,195-205); it is optional here: used when importable, else this plain-text table with the same
five calls (add_column / __setitem__ / update / next_row / close)."""
import sys

try:  # pragma: no cover - not installed in the build image
    from progress_table import ProgressTable as _ProgressTable
except ImportError:
    _ProgressTable = None


class PlainTable:
    def __init__(self, file=None, **kwargs):
        self.file = file if file is not None else sys.stdout
        self.columns = []
        self.rows = []
        self._row = {}
        self._header_printed = False

    def add_column(self, name, **kwargs):
        if name not in self.columns:
            self.columns.append(name)

    def __setitem__(self, name, value):
        self.add_column(name)
        self._row[name] = value

    def __getitem__(self, name):
        return self._row.get(name)

    def update(self, name, value, **kwargs):
        self[name] = value

    @staticmethod
    def _fmt(v):
        if v is None:
            return ''
        if hasattr(v, 'item') and getattr(v, 'numel', lambda: 2)() == 1:
            v = v.item()
        if isinstance(v, float):
            return f'{v:.4f}'
        return str(v)

    def next_row(self, **kwargs):
        if not self._header_printed:
            print(' | '.join(f'{c:>14}' for c in self.columns), file=self.file)
            self._header_printed = True
        print(' | '.join(f'{self._fmt(self._row.get(c)):>14}' for c in self.columns), file=self.file)
        self.rows.append(self._row)
        self._row = {}

    def close(self):
        if self._row:
            self.next_row()
        flush = getattr(self.file, 'flush', None)
        if flush:
            flush()


ProgressTable = _ProgressTable or PlainTable


class EpochTable:
    """The per-epoch progress table of a Stage: parses the stage's column spec once, then fills one row per epoch
    from the MetricTracker.  A column spec entry is a metric name, or a dict with 'name' (display), 'metric'
    (tracker name, None = filled by the caller) and optional extra keys handed to the table backend."""

    def __init__(self, spec, sink):
        self.columns = [self._parse(entry) for entry in spec]
        self.backend = ProgressTable(file=sink)
        for col in self.columns:
            extra = {k: v for k, v in col.items() if k not in ('name', 'metric')}
            self.backend.add_column(col['name'], **extra)

    @staticmethod
    def _parse(entry):
        if isinstance(entry, str):
            return {'name': entry, 'metric': entry}
        if not isinstance(entry, dict):
            raise ValueError(f'Invalid column: {entry}. Must be a string or a dict.')
        for key in ('name', 'metric'):
            if key not in entry:
                raise ValueError(f'Column dict must contain a "{key}" key')
        return dict(entry)

    def has(self, display_name):
        return any(c['name'] == display_name for c in self.columns)

    def set(self, display_name, value):
        self.backend.update(display_name, value)

    def __setitem__(self, display_name, value):
        self.backend[display_name] = value

    # the reference hands stages the ProgressTable itself (stage.py:147): user code written against it keeps working
    def __getitem__(self, display_name):
        return self.backend[display_name]

    def update(self, display_name, value, **kwargs):
        self.backend.update(display_name, value, **kwargs)

    def add_column(self, display_name, **kwargs):
        if not self.has(display_name):
            self.columns.append({'name': display_name, 'metric': None})
        self.backend.add_column(display_name, **kwargs)

    def next_row(self, **kwargs):
        self.backend.next_row(**kwargs)

    def emit_row(self, tracker):
        for col in self.columns:
            if col['metric'] is not None:
                self.backend.update(col['name'], tracker[col['metric']][-1])
        self.backend.next_row()

    def close(self):
        self.backend.close()
