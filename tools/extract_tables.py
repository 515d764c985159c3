#!/usr/bin/env python
"""Extract the numeric DATA tables of the SBDART reference into one .npz bundle.

The reference keeps ~14k lines of physical tables (band-model coefficients,
continua, Mie tables, solar spectra, standard atmospheres, albedos) in Fortran
DATA statements.  They are DATA, not code: this tool parses the fixed-form
sources under /root/reference and writes
    sbdart_b200/frontend/tables.npz
keyed "<file>/<unit>/<name>" (unit = enclosing subroutine / function / module).
Literals without a D exponent or _kr suffix are default REAL in the reference
and are therefore rounded through float32 before widening.

Run in the build container only (the GPU box has no /root/reference); the
generated bundle is committed.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

REF = "/root/reference"
FILES = ["taugas.f", "atms.f", "spectra.f", "taucloud.f", "tauaero.f", "disutil.f", "params.f"]


def logical_lines(path):
    """Join fixed-form continuation lines; drop comments."""
    out, cur = [], None
    for raw in open(path, errors="replace"):
        line = raw.rstrip("\n")
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        # strip trailing ! comment (no string literals in the table statements we keep)
        if "!" in line and "'" not in line and '"' not in line:
            line = line[: line.index("!")]
        if len(line) > 5 and line[5] not in " 0" and line[:5].strip() == "":
            if cur is not None:
                cur += line[6:]
            continue
        if cur is not None:
            out.append(cur)
        cur = line[6:] if len(line) > 6 else ""
        lab = line[:5].strip()
        if lab and not lab.isdigit():
            cur = line  # free-ish form line (module-level statements start in col 7 anyway)
    if cur is not None:
        out.append(cur)
    return out


NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?(_\w+)?$")


def parse_value(tok):
    tok = tok.strip()
    if tok.lower() in (".true.", ".t."):
        return 1.0, True
    if tok.lower() in (".false.", ".f."):
        return 0.0, True
    m = NUM.match(tok)
    if not m:
        raise ValueError(tok)
    is_int = re.match(r"^[+-]?\d+$", tok) is not None
    dbl = ("d" in tok.lower()) or (m.group(3) is not None)
    core = tok
    if m.group(3):
        core = tok[: tok.rindex("_")]
    v = float(core.lower().replace("d", "e"))
    if not dbl and not is_int:
        v = float(np.float32(v))
    return v, is_int


def split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def parse_values(vs, params):
    vals, ints = [], True
    for tok in split_top(vs):
        tok = tok.strip()
        if not tok:
            continue
        if "*" in tok:
            r, v = tok.split("*", 1)
            r = r.strip()
            rep = int(params.get(r.lower(), r)) if not r.isdigit() else int(r)
            v = v.strip()
            if v.lower() in params:
                val, isint = float(params[v.lower()]), False
            else:
                val, isint = parse_value(v)
            vals += [val] * rep
            ints = ints and isint
        else:
            if tok.lower() in params:
                val, isint = float(params[tok.lower()]), False
            else:
                val, isint = parse_value(tok)
            vals.append(val)
            ints = ints and isint
    return vals, ints


def eval_int(expr, params):
    expr = expr.strip().lower()
    try:
        return int(eval(expr, {"__builtins__": {}}, {k: int(v) if float(v).is_integer() else v for k, v in params.items()}))
    except Exception:
        return None


def extract_file(path):
    lines = logical_lines(path)
    tables = {}
    unit, params = "main", {}
    shapes = {}
    for ln in lines:
        s = ln.strip()
        low = s.lower()
        m = re.match(r"^(?:real\(kr\)\s+|real\s+|integer\s+|logical\s+)?(subroutine|function|module|program)\s+(\w+)", low)
        if m and not low.startswith("end") and not low.startswith("module procedure"):
            if m.group(1) == "module" and low.startswith("module"):
                unit, params, shapes = m.group(2), {}, {}
                continue
            if m.group(1) != "module":
                unit, params, shapes = m.group(2), dict(params) if False else {}, {}
                continue
        # integer/real parameters (simple constants only)
        pm = re.match(r"^(integer|real\(kr\)|real)\s*,\s*parameter\s*::\s*(.*)$", low)
        if pm:
            for item in split_top(pm.group(2)):
                if "=" in item:
                    k, v = item.split("=", 1)
                    try:
                        params[k.strip()] = parse_value(v.strip())[0]
                    except ValueError:
                        ev = eval_int(v, params)
                        if ev is not None:
                            params[k.strip()] = ev
            continue
        pm = re.match(r"^parameter\s*\((.*)\)\s*$", low)
        if pm:
            for item in split_top(pm.group(1)):
                if "=" in item:
                    k, v = item.split("=", 1)
                    try:
                        params[k.strip()] = parse_value(v.strip())[0]
                    except ValueError:
                        ev = eval_int(v, params)
                        if ev is not None:
                            params[k.strip()] = ev
            continue
        # declarations with dimensions -> shapes
        dm = re.match(r"^(real\(kr\)|real|integer|logical|dimension|double precision)\s*(?:,\s*save\s*)?(?:::)?\s*(.*)$", low)
        if dm and not low.startswith("data"):
            for item in split_top(dm.group(2)):
                im = re.match(r"^\s*(\w+)\s*\(([^)]*)\)", item)
                if im:
                    dims = []
                    ok = True
                    for d in im.group(2).split(","):
                        if ":" in d:
                            lo, hi = d.split(":")
                            lo, hi = eval_int(lo, params), eval_int(hi, params)
                            if lo is None or hi is None:
                                ok = False
                                break
                            dims.append((lo, hi))
                        else:
                            hi = eval_int(d, params)
                            if hi is None:
                                ok = False
                                break
                            dims.append((1, hi))
                    if ok:
                        shapes[im.group(1)] = dims
            # module-level initialisers  real(kr) :: x(3) = (/ ... /)
            continue
        if not low.startswith("data"):
            continue
        body = s[4:].strip()
        # groups: objlist / values / [,] objlist / values / ...
        pos = 0
        while pos < len(body):
            i1 = body.find("/", pos)
            if i1 < 0:
                break
            i2 = body.find("/", i1 + 1)
            if i2 < 0:
                break
            objs = body[pos:i1].strip().lstrip(",").strip()
            try:
                vals, isint = parse_values(body[i1 + 1: i2], params)
            except ValueError as e:
                print(f"  skip {os.path.basename(path)}:{unit}: {objs[:40]} ({e})", file=sys.stderr)
                pos = i2 + 1
                continue
            pos = i2 + 1
            vi = 0
            for obj in split_top(objs):
                obj = obj.strip()
                if not obj:
                    continue
                # implied do: (name(i),i=lo,hi)
                im = re.match(r"^\(\s*(\w+)\s*\(\s*(\w+)\s*\)\s*,\s*(\w+)\s*=\s*([^,]+),\s*([^,)]+)\)$", obj)
                rm = re.match(r"^(\w+)\s*\(\s*([^:()]+)\s*:\s*([^:()]+)\s*\)$", obj)
                sm = re.match(r"^(\w+)\s*\(\s*([^:(),]+)\s*\)$", obj)
                im2 = re.match(r"^\(\s*(\w+)\s*\(\s*(\w+)\s*,\s*(\w+)\s*\)\s*,\s*(\w+)\s*=\s*([^,]+),\s*([^,)]+)\)$", obj)
                if im2 and im2.group(1).lower() in shapes and len(shapes[im2.group(1).lower()]) == 2:
                    # (a(i,J), i=lo,hi) or (a(J,i), i=lo,hi): one column / row of a 2-D table
                    name = im2.group(1).lower()
                    var = im2.group(4).lower()
                    l0, h0 = eval_int(im2.group(5), params), eval_int(im2.group(6), params)
                    d1 = shapes[name][0][1] - shapes[name][0][0] + 1
                    n = h0 - l0 + 1
                    chunk = vals[vi: vi + n]
                    vi += n
                    key = f"{os.path.basename(path)[:-2]}/{unit}/{name}"
                    ent = tables.setdefault(key, {"base": 1, "vals": {}, "int": True, "shape": shapes.get(name)})
                    ent["int"] = ent["int"] and isint
                    for k, v in enumerate(chunk):
                        if im2.group(2).lower() == var:
                            i, j = l0 + k, eval_int(im2.group(3), params)
                        else:
                            i, j = eval_int(im2.group(2), params), l0 + k
                        ent["vals"][(j - 1) * d1 + i] = v
                    continue
                if im:
                    name, lo, hi = im.group(1).lower(), eval_int(im.group(4), params), eval_int(im.group(5), params)
                elif rm:
                    name, lo, hi = rm.group(1).lower(), eval_int(rm.group(2), params), eval_int(rm.group(3), params)
                elif sm:
                    name = sm.group(1).lower()
                    lo = hi = eval_int(sm.group(2), params)
                elif re.match(r"^\w+$", obj):
                    name = obj.lower()
                    if name in shapes:
                        n = 1
                        for a, b in shapes[name]:
                            n *= b - a + 1
                        lo, hi = shapes[name][0][0], shapes[name][0][0] + n - 1
                    else:
                        # scalar, or array whose size is the remaining value count
                        nobj = len([o for o in split_top(objs) if o.strip()])
                        n = len(vals) - vi if nobj == 1 else 1
                        lo, hi = 1, n
                else:
                    print(f"  skip object {obj[:50]} in {unit}", file=sys.stderr)
                    continue
                if lo is None or hi is None:
                    print(f"  skip object {obj[:50]} in {unit} (bounds)", file=sys.stderr)
                    continue
                n = hi - lo + 1
                chunk = vals[vi: vi + n]
                vi += n
                key = f"{os.path.basename(path)[:-2]}/{unit}/{name}"
                base = shapes[name][0][0] if name in shapes else 1
                ent = tables.setdefault(key, {"base": base, "vals": {}, "int": True, "shape": shapes.get(name)})
                ent["int"] = ent["int"] and isint
                for k, v in enumerate(chunk):
                    ent["vals"][lo + k] = v
    out = {}
    for key, ent in tables.items():
        idx = sorted(ent["vals"])
        lo, hi = idx[0], idx[-1]
        base = min(ent["base"], lo)
        arr = np.zeros(hi - base + 1)
        for k, v in ent["vals"].items():
            arr[k - base] = v
        if ent["shape"] and len(ent["shape"]) > 1:
            dims = [b - a + 1 for a, b in ent["shape"]]
            if int(np.prod(dims)) == arr.size:
                arr = arr.reshape(dims[::-1]).T      # Fortran column-major -> [i][j]
        out[key] = arr.astype(np.int64) if ent["int"] else arr
    return out


def extract_extras():
    """Tables the generic DATA parser does not see: array-constructor initialisers
    (`x = (/ ... /)`), DATA statements on array sections (`data aerstr(1:naerw,k,j) /../`)
    and the wavelength limits of the sensor filters (named PARAMETERs)."""
    out = {}

    def f32s(txt):
        return np.array([float(np.float32(float(t))) for t in re.split(r"[\s,]+", txt.strip()) if t])

    # tauaero.f: awl (aeroblk), alt / aden05 / aden23 (aerzstd), aerstr (aestrat)
    text = " ".join(logical_lines(os.path.join(REF, "tauaero.f")))
    for unit, name in (("aeroblk", "awl"), ("aerzstd", "alt"), ("aerzstd", "aden05"), ("aerzstd", "aden23")):
        m = re.search(r"\b" + name + r"\s*=\s*\(/(.*?)/\)", text, re.S)
        out[f"tauaero/{unit}/{name}"] = f32s(m.group(1))
    aer = np.zeros((47, 3, 4))
    for m in re.finditer(r"data\s+aerstr\(1:naerw,(\d),(\d)\)\s*/(.*?)/", text, re.S | re.I):
        v = f32s(m.group(3))
        assert v.size == 47, (m.group(1), m.group(2), v.size)
        aer[:, int(m.group(1)) - 1, int(m.group(2)) - 1] = v
    out["tauaero/aestrat/aerstr"] = aer
    # spectra.f: wmn / wmx of every sensor filter routine (spectra.f:3420-4430)
    unit = None
    for ln in logical_lines(os.path.join(REF, "spectra.f")):
        low = ln.strip().lower()
        m = re.match(r"^subroutine\s+(\w+)\s*\(srr,wmin,wmax,nnf\)", low)
        if m:
            unit = m.group(1)
            continue
        m = re.match(r"^real\(kr\),\s*parameter\s*::\s*wmn=([\d.]+)\s*,\s*wmx=([\d.]+)", low)
        if m and unit:
            out[f"spectra/{unit}/wmn"] = np.array(float(np.float32(float(m.group(1)))))
            out[f"spectra/{unit}/wmx"] = np.array(float(np.float32(float(m.group(2)))))
            unit = None
    return out


def main():
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "sbdart_b200", "frontend", "tables.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    allt = {}
    for f in FILES:
        t = extract_file(os.path.join(REF, f))
        print(f, len(t), "tables,", sum(v.size for v in t.values()), "values")
        allt.update(t)
    ex = extract_extras()
    print("extras", len(ex), "tables")
    allt.update(ex)
    np.savez_compressed(dst, **allt)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
