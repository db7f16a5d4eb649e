"""Scalar logging with the reference's call contract `log_value(name, value, step)`
(tb_json_logger.py:38-84): values are kept in memory as {step: {name: value}} and exported to
result.json.  tensorboard_logger is optional here (absent in this image -> json only)."""
import json
import os
from collections import OrderedDict

from utils import check_dir_exists

try:                                     # pragma: no cover - optional dependency
    import tensorboard_logger as _tb
except Exception:                        # noqa: BLE001
    _tb = None

_values = OrderedDict()
_tb_logger = None


def configure(tb_path=None, json_path=None, resume=False):
    """Start a run; optionally reload a previous result.json (cfg.resume_result_json)."""
    global _tb_logger
    _values.clear()
    if resume and json_path and os.path.isfile(json_path):
        with open(json_path) as fh:
            for k, v in json.load(fh).items():
                _values[int(k)] = v
    if _tb is not None and tb_path:
        _tb_logger = _tb.Logger(tb_path, flush_secs=5)


def log_value(name, value, step):
    _values.setdefault(int(step), OrderedDict())[name] = float(value)
    if _tb_logger is not None:
        _tb_logger.log_value(name, value, step)


def get_values():
    return _values


def export_to_json(fn, it_filter=None):
    check_dir_exists(fn)
    keep = {it: v for it, v in _values.items() if it_filter is None or it_filter(it)}
    with open(fn, 'w') as fh:
        json.dump(keep, fh, indent=2, sort_keys=True)
