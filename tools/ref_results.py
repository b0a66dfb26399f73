"""Decode the reference's published Monte-Carlo curves (``results/*_results``: what ``scripts/evaluate.py:204-207`` saves —
[Eb/N0 grid, {key: BER}, {key: BLER}(, {key: extra})] with key = (receiver name, number of active UEs, MCS index(, sweep
value)) — pickled with one TensorFlow symbol inside, which is stubbed out here) into a small JSON file, and print a curve.

    python tools/ref_results.py dump  [/root/reference/results]      -> tests/golden/ref_published_curves.json
    python tools/ref_results.py show nrx_large_results ["Neural Receiver" 2 0]

The JSON holds published NUMBERS (the perf / quality baselines a BLER harness is compared against, like BASELINE.md),
no code.  ``tools/bler_sweep.py --published nrx_large_results`` prints the published BLER of the 2-UE neural receiver
next to its own column for orientation (not a parity claim while the LDPC code is the stand-in)."""
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "ref_published_curves.json")


class _Tensor:
    """Stands in for ``tf.convert_to_tensor(ndarray)`` inside the pickles: keeps the array."""

    def __init__(self, *args, **kwargs):
        self.value = args[0] if args else None


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("numpy", "builtins", "collections", "_codecs"):
            return super().find_class(module, name)
        return _Tensor


def _arr(v):
    v = v.value if isinstance(v, _Tensor) else v
    return [float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)]


def decode(path):
    with open(path, "rb") as f:
        x = _Unpickler(f).load()
    out = {"ebno_db": _arr(x[0]), "ber": {}, "bler": {}}
    for name, d in (("ber", x[1]), ("bler", x[2])):
        for k, v in d.items():
            out[name]["|".join(str(int(e)) if not isinstance(e, str) else e for e in k)] = _arr(v)
    return out


def load_published(path=OUT):
    with open(path) as f:
        return json.load(f)


def published_at(name, key, ebno_db, kind="bler"):
    """Published value of curve `key` of results file `name` at `ebno_db` (None when the point or the curve is absent)."""
    try:
        res = load_published()[name]
        curve = res[kind][key]
    except (OSError, KeyError):
        return None
    for e, v in zip(res["ebno_db"], curve):
        if abs(e - ebno_db) < 1e-6:
            return v
    return None


def main():
    if len(sys.argv) < 2 or sys.argv[1] not in ("dump", "show"):
        raise SystemExit(__doc__)
    if sys.argv[1] == "dump":
        d = sys.argv[2] if len(sys.argv) > 2 else "/root/reference/results"
        res = {f: decode(os.path.join(d, f)) for f in sorted(os.listdir(d))}
        with open(OUT, "w") as f:
            json.dump(res, f, indent=0, sort_keys=True)
        print(f"{OUT}: {len(res)} files, {sum(len(v['bler']) for v in res.values())} BLER curves")
        return
    res = load_published()[sys.argv[2]]
    want = "|".join(sys.argv[3:]) if len(sys.argv) > 3 else None
    print("# ebno_db " + " ".join(f"{e:9.1f}" for e in res["ebno_db"]))
    for k, v in res["bler"].items():
        if want is None or k == want:
            print(f"BLER {k:40s} " + " ".join(f"{x:9.2e}" for x in v))


if __name__ == "__main__":
    main()
