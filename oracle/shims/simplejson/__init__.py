"""Test-only stand-in for `simplejson` (absent from this image): forwards to the stdlib json."""
import json as _json

JSONDecodeError = _json.JSONDecodeError
loads = _json.loads


def load(fp, **kw):
    return _json.load(fp, **kw)


def dumps(obj, **kw):
    kw.pop('ignore_nan', None)
    return _json.dumps(obj, **kw)


def dump(obj, fp, **kw):
    kw.pop('ignore_nan', None)
    return _json.dump(obj, fp, **kw)
