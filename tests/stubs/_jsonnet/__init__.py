"""Stand-in for the `_jsonnet` module (not installed in this image; test infrastructure): evaluates the small subset of
Jsonnet that robogym's rearrange material files use (robogym/envs/rearrange/materials/*.jsonnet, loaded by
robogym/envs/rearrange/common/utils.py:1041-1045) -- one object literal with bare keys, strings and numbers, `#` comments and
trailing commas, optionally `(import "base.libsonnet") + { ... }` where `key+:` merges into the imported object's field."""
import json
import os
import re


def _parse_object(text):
    """object literal -> (dict, set of keys written with `+:`), nested objects handled recursively"""
    text = re.sub(r"#[^\n]*", "", text)
    plus = set()

    def key(mo):
        if mo.group(2):
            plus.add(mo.group(1))
        return '"%s":' % mo.group(1)

    js = re.sub(r"([A-Za-z_][A-Za-z_0-9]*)\s*(\+?):", key, text)
    js = re.sub(r",\s*([}\]])", r"\1", js)
    return json.loads(js), plus


def _merge(base, over, plus):
    out = dict(base)
    for k, v in over.items():
        if k in plus and isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = {**out[k], **v}
        else:
            out[k] = v
    return out


def evaluate_file(path):
    text = open(path).read()
    mo = re.match(r'\s*\(\s*import\s+"([^"]+)"\s*\)\s*\+\s*', text)
    if mo:
        base = json.loads(evaluate_file(os.path.join(os.path.dirname(path), mo.group(1))))
        over, plus = _parse_object(text[mo.end():])
        return json.dumps(_merge(base, over, plus))
    obj, _ = _parse_object(text)
    return json.dumps(obj)
