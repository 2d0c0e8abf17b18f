#!/usr/bin/env python
"""Provenance of tests/golden/reference_specs.json.

Re-extracts, from the reference's own test sources under /root/reference (present in the build container, absent on
the GPU box), the inputs and expected outputs this repo pins its oracle against, recomputes the "derived" block from
the reference's formulas in 50-digit arithmetic (independently of oracle/ref_oracle.py), and compares the result with
the committed JSON:

    python tests/golden/extract_reference_specs.py            # prints OK or the differences, exit status 0 / 1
    python tests/golden/extract_reference_specs.py --write    # regenerate the JSON

Only numbers are read (Vector(...) literals and tolerances); no reference source text is copied.
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

import mpmath as mp

REF = Path("/root/reference/src/test/scala/thylacine/model/components")
OUT = Path(__file__).resolve().parent / "reference_specs.json"


def _num(tok: str) -> float:
    tok = tok.strip().rstrip("d")
    return float(tok if not tok.startswith(".") and not tok.startswith("-.") else tok.replace(".", "0.", 1))


def _vector(text: str, start: int):
    """Parse the Vector(...) literal that begins at text[start] == 'V' into (nested) python lists."""
    assert text.startswith("Vector(", start), text[start:start + 20]
    i, depth, items, cur, j = start + len("Vector("), 1, [], "", start + len("Vector(")
    while depth:
        c = text[j]
        if text.startswith("Vector(", j):
            sub, j = _vector(text, j)
            items.append(sub)
            cur = ""
            continue
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                if cur.strip():
                    items.append(_num(cur))
                break
        elif c == "," and depth == 1:
            if cur.strip():
                items.append(_num(cur))
            cur = ""
            j += 1
            continue
        cur += c if depth >= 1 and c not in "()" else ""
        j += 1
    return items, j + 1


def _named_vector(block: str, name: str):
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*(?=Vector\()", block)
    assert m, name
    return _vector(block, m.end())[0]


def _block(text: str, start_pat: str) -> str:
    m = re.search(start_pat, text)
    assert m, start_pat
    depth, j = 0, text.index("(", m.end() - 1) if text[m.end() - 1] != "(" else m.end() - 1
    k = j
    while True:
        if text[k] == "(":
            depth += 1
        elif text[k] == ")":
            depth -= 1
            if depth == 0:
                return text[j:k + 1]
        k += 1


def extract():
    fx = (REF / "ComponentFixture.scala").read_text()
    out = {"_source": json.loads(OUT.read_text())["_source"] if OUT.exists() else ""}

    spec = (REF / "prior" / "GaussianPriorSpec.scala").read_text()
    grad_case = spec[spec.index("generate the correct gradient of the logPdf"):]
    gb = _block(grad_case, r"fromConfidenceIntervals\[IO\]\(")
    m = re.search(r"rawLogPdfGradientAt\((?=Vector\()", grad_case)
    x = _vector(grad_case, m.end())[0]
    exp = re.search(r"shouldBe Vector\(([^\n]*)\)\s*$", grad_case[m.end():], re.M).group(1)
    expected = [e.strip() for e in exp.split(",")]
    assert expected == ["-.25", "4.0 / 9", "-1"], expected          # kept symbolic in the JSON: -0.25, 4/9, -1
    out["gaussian_prior_gradient"] = {"mean": _named_vector(gb, "values"), "ci": _named_vector(gb, "confidenceIntervals"),
                                      "x": x, "expected": ["-0.25", "4/9", "-1"]}

    def prior(val_name, label_name):
        b = _block(fx, r"val " + val_name + r":[^=]*=\s*\n?\s*GaussianPrior\.fromConfidenceIntervals\[IO\]\(")
        label = re.search(r'val ' + label_name + r': String\s*=\s*"([^"]+)"', fx).group(1)
        return {"label": label, "mean": _named_vector(b, "values"), "ci": _named_vector(b, "confidenceIntervals")}

    def uniform(val_name, label_name):
        b = _block(fx, r"val " + val_name + r":[^=]*=\s*\n?\s*UniformPrior\.fromBounds\[IO\]\(")
        label = re.search(r'val ' + label_name + r': String\s*=\s*"([^"]+)"', fx).group(1)
        return {"label": label, "max": _named_vector(b, "maxBounds"), "min": _named_vector(b, "minBounds")}

    def likelihood(def_name):
        b = _block(fx, r"def " + def_name + r"\(implicit[^=]*=\s*GaussianLinearLikelihood\.of\[IO\]\(")
        return {"A": _named_vector(b, "coefficients"), "y": _named_vector(b, "measurements"),
                "ci": _named_vector(b, "uncertainties")}

    out["foo_prior"] = prior("fooPrior", "fooPriorLabel")
    out["foo_uniform_prior"] = uniform("fooUniformPrior", "fooUniformPriorLabel")
    foo = likelihood("fooLikelihoodF")
    lin = (REF / "likelihood" / "GaussianLinearLikelihoodSpec.scala").read_text()
    g = [_vector(lin, m.end())[0] for m in re.finditer(r'shouldBe Map\("foo" -> (?=Vector\()', lin)]
    foo["grad_at_1_2"], foo["grad_at_3_2"] = g[0], g[1]
    foo["fd_differential"] = _num(re.search(r"differential\s*=\s*([.\d]+)", fx).group(1))
    nl = (REF / "likelihood" / "GaussianLikelihoodSpec.scala").read_text()
    foo["fd_abs_tol"] = float(re.search(r"shouldBe \(0d \+- ([\de.-]+)\)", nl).group(1))
    out["foo_likelihood"] = foo
    out["bar_prior"] = prior("barPrior", "barPriorLabel")
    out["bar_uniform_prior"] = uniform("barUniformPrior", "barUniformPriorLabel")
    out["bar_likelihood"] = likelihood("barLikelihoodF")
    out["derived"] = derived(out)
    return out


def derived(spec):
    """The reference's formulas (GaussianDistribution.scala:52-61 via Breeze's MultivariateGaussian.logPdf,
    CauchyDistribution.scala:51-79) at the fixture inputs, 50 digits."""
    mp.mp.dps = 50
    A = [[mp.mpf(v) for v in row] for row in spec["foo_likelihood"]["A"]]
    y = [mp.mpf(v) for v in spec["foo_likelihood"]["y"]]
    var = [(mp.mpf(str(c)) / 2) ** 2 for c in spec["foo_likelihood"]["ci"]]

    def resid(x):
        return [y[i] - sum(A[i][k] * x[k] for k in range(2)) for i in range(2)]

    def gauss(r, v):
        n = len(r)
        return -sum(ri * ri / vi for ri, vi in zip(r, v)) / 2 - (mp.mpf(n) / 2 * mp.log(2 * mp.pi) + sum(mp.log(mp.sqrt(vi)) for vi in v))

    pm = [mp.mpf(v) for v in spec["foo_prior"]["mean"]]
    pv = [(mp.mpf(c) / 2) ** 2 for c in spec["foo_prior"]["ci"]]
    d = {"foo_likelihood_logpdf_at_1_2": gauss(resid([1, 2]), var), "foo_likelihood_logpdf_at_3_2": gauss(resid([3, 2]), var),
         "foo_prior_logpdf_at_1_2": gauss([pm[0] - 1, pm[1] - 2], pv), "foo_prior_logpdf_at_3_2": gauss([pm[0] - 3, pm[1] - 2], pv)}
    r = resid([3, 2])
    n = 2
    q = sum(ri * ri / vi for ri, vi in zip(r, var))
    d["foo_cauchy_logpdf_at_3_2"] = (mp.log(mp.gamma(mp.mpf(1 + n) / 2)) - mp.log(mp.gamma(mp.mpf(1) / 2)) - mp.mpf(n) / 2 * mp.log(mp.pi)
                                     - mp.log(var[0] * var[1]) / 2 - mp.mpf(1 + n) / 2 * mp.log(1 + q))
    # the reference's "gradient" keeps the + sign (CauchyDistribution.scala:73-78), w.r.t. the observation;
    # through the linear forward model: A^T (mult * Sigma^-1 (f - y))
    mult = mp.mpf(1 + n) / (1 + q)
    gf = [mult * (-r[i]) / var[i] for i in range(2)]
    d["foo_cauchy_reference_gradient_at_3_2"] = [sum(A[i][k] * gf[i] for i in range(2)) for k in range(2)]
    return {k: ([float(mp.nstr(x, 9)) for x in v] if isinstance(v, list) else float(v)) for k, v in d.items()}


def main():
    got = extract()
    if "--write" in sys.argv:
        OUT.write_text(json.dumps(got, indent=2) + "\n")
        print("written", OUT)
        return 0
    want = json.loads(OUT.read_text())
    bad = []
    for k in want:
        if k == "_source":
            continue
        if k == "derived":
            for dk, dv in want[k].items():
                gv = got[k][dk]
                pairs = list(zip(dv, gv)) if isinstance(dv, list) else [(dv, gv)]
                for a, b in pairs:
                    if abs(a - b) > 1e-8 * max(1.0, abs(a)) and not (isinstance(dv, list) and abs(a - b) < 1e-7):
                        bad.append((dk, dv, gv))
        elif want[k] != got[k]:
            bad.append((k, want[k], got[k]))
    if bad:
        for b in bad:
            print("DIFFERS:", b)
        return 1
    print("OK: tests/golden/reference_specs.json matches the reference's test sources and formulas")
    return 0


if __name__ == "__main__":
    sys.exit(main())
