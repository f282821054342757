"""Per-car state record layout, read from ``include/ncg_b200.h`` so Python and CUDA cannot drift."""
from __future__ import annotations

import os
import re
from typing import Dict

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "ncg_b200.h")


def _parse_enum(text: str, name: str, env: Dict[str, int]) -> Dict[str, int]:
    m = re.search(r"enum\s+" + name + r"\s*\{(.*?)\};", text, re.S)
    if not m:
        raise RuntimeError(f"enum {name} not found in ncg_b200.h")
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out: Dict[str, int] = {}
    nxt = 0
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, expr = [s.strip() for s in item.split("=", 1)]
            expr = expr.replace("u", "") if re.fullmatch(r"[0-9u<\s]+", expr) else expr
            val = int(eval(expr, {"__builtins__": {}}, {**env, **out}))
        else:
            k, val = item, nxt
        out[k] = val
        nxt = val + 1
    return out


def _load():
    with open(_HEADER) as f:
        text = f.read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(NCG_[A-Z_]+)\s+(\d+)", text)}
    rec = _parse_enum(text, "NcgRecordField", defs)
    flags = _parse_enum(text, "NcgFlagBits", defs)
    return defs, rec, flags


DEFS, R, F = _load()
RECORD_WORDS = DEFS["NCG_RECORD_WORDS"]
OBS_DIM = DEFS["NCG_OBS_DIM"]
MAX_CONTACTS = DEFS["NCG_MAX_CONTACTS"]
MAX_TOUCHING = DEFS["NCG_MAX_TOUCHING"]
MAX_ACTIVE = DEFS["NCG_MAX_ACTIVE"]
VEL_HISTORY = DEFS["NCG_VEL_HISTORY"]
