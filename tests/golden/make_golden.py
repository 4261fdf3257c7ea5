#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the Hex8 assembly path.

Run in the build container only (needs /root/reference, which does not travel
to the GPU box):

    python tests/golden/make_golden.py

It parses the `static constexpr double name[..] = {...}` tables out of the
reference unit tests and writes them, unmodified, as JSON fixtures next to this
script.  Nothing is computed here: every number in tests/golden/*.json is a
literal from the cited reference file.

Sources (relative to /root/reference):
  unit_tests/kernels/UnitTestMomentumAdvDiffElem.C:17-269   momentum adv-diff lhs/rhs
  unit_tests/kernels/UnitTestContinuityAdvElem.C:18-57      continuity default / reduced-sens
  unit_tests/kernels/UnitTestScalarAdvDiffElem.C:18-31      scalar adv-diff (perturbed element)
  unit_tests/kernels/UnitTestScalarDiffElem.C:18-28         scalar diffusion (Laplacian stencil)
  unit_tests/kernels/UnitTestMomentumMassElem.C:17-140      momentum mass consistent / lumped
  unit_tests/UnitTestKokkosMEGold.h:13-79                   SCV grad-op tables (perturbed element)
  unit_tests/UnitTestTpetra.C:86-114                        2-element assembled matrix
  unit_tests/kernels/UnitTestSSTSrcElem.C:17-176            SST tke / tke-DES / sdr source lhs/rhs
  unit_tests/kernels/UnitTestMomentumNSOElem.C:14-316       momentum NSO (2nd order; pde and linearized residual)
"""
import json
import os
import re
import sys

REF = "/root/reference/unit_tests"
HERE = os.path.dirname(os.path.abspath(__file__))

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def _eval_token(tok):
    # tokens like "2.0*2" or "-0.8*2" appear in UnitTestTpetra.C
    tok = tok.strip()
    if "*" in tok:
        a, b = tok.split("*")
        return float(a) * float(b)
    return float(tok)


def parse_tables(path):
    """Return {qualified_name: flat list of doubles} for every constexpr double table."""
    src = open(path).read()
    # strip comments
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    # track namespaces by a simple brace walk
    ns_stack = []
    brace_kinds = []
    i = 0
    n = len(src)
    tok_ns = re.compile(r"namespace\s+(\w+)?\s*\{")
    tok_tab = re.compile(r"static\s+(?:constexpr|const)\s+double\s+(\w+)\s*((?:\[\s*\d*\s*\])+)\s*=\s*\{")
    while i < n:
        m_ns = tok_ns.match(src, i)
        if m_ns:
            ns_stack.append(m_ns.group(1) or "")
            brace_kinds.append("ns")
            i = m_ns.end()
            continue
        m_tab = tok_tab.match(src, i)
        if m_tab:
            # find the matching closing brace of the initialiser
            depth = 1
            j = m_tab.end()
            while depth:
                c = src[j]
                if c == "{":
                    depth += 1
                elif c == "}":
                    depth -= 1
                j += 1
            body = src[m_tab.end():j - 1]
            toks = re.findall(NUM + r"(?:\s*\*\s*" + NUM + r")?", body)
            vals = [_eval_token(t) for t in toks]
            dims = [int(d) for d in re.findall(r"\[\s*(\d+)\s*\]", m_tab.group(2))]
            cnt = 1
            for d in dims:
                cnt *= d
            assert len(vals) == cnt, (path, m_tab.group(1), len(vals), cnt)
            qn = "::".join([s for s in ns_stack if s] + [m_tab.group(1)])
            out[qn] = {"dims": dims, "values": vals}
            i = j
            continue
        c = src[i]
        if c == "{":
            brace_kinds.append("other")
        elif c == "}":
            if brace_kinds:
                k = brace_kinds.pop()
                if k == "ns":
                    ns_stack.pop()
        i += 1
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    files = {
        "momentum_advdiff": "kernels/UnitTestMomentumAdvDiffElem.C",
        "continuity_adv": "kernels/UnitTestContinuityAdvElem.C",
        "scalar_advdiff": "kernels/UnitTestScalarAdvDiffElem.C",
        "scalar_diff": "kernels/UnitTestScalarDiffElem.C",
        "momentum_mass": "kernels/UnitTestMomentumMassElem.C",
        "kokkos_me_gold": "UnitTestKokkosMEGold.h",
        "tpetra_2elem": "UnitTestTpetra.C",
        "sst_src": "kernels/UnitTestSSTSrcElem.C",
        "momentum_nso": "kernels/UnitTestMomentumNSOElem.C",
    }
    for key, rel in files.items():
        tabs = parse_tables(os.path.join(REF, rel))
        assert tabs, rel
        doc = {"source": "unit_tests/" + rel, "tables": tabs}
        with open(os.path.join(HERE, key + ".json"), "w") as f:
            json.dump(doc, f, indent=0, separators=(",", ":"))
        print(key, {k: v["dims"] for k, v in tabs.items()})


if __name__ == "__main__":
    main()
