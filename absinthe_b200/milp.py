"""A small mixed-integer linear modelling layer: build, write / read CPLEX-LP text, solve with HiGHS.

The reference writes its model as LP text and shells out to the ``cplex`` interactive optimiser
(ref stencil_optimizer.py:762-808), then reads CPLEX's XML solution (ref :826-830, :899).  Here the
model is an object (``Model``): the optimiser (``optimizer.py``) adds variables and rows to it, it can
be solved in-process (``scipy.optimize.milp`` = HiGHS), and it still round-trips through the two file
formats -- ``write_lp`` / ``read_lp`` and ``write_solution`` / ``read_solution`` -- so that a real
``cplex`` (or ``tools/cplex``, a command-line front end of this module) stays interchangeable.
"""

import os
import re
from xml.dom.minidom import parse
from xml.sax.saxutils import quoteattr

CONTINUOUS, GENERAL, BINARY = 0, 1, 2
_SENSES = ("<=", ">=", "=")


class Model:
    """minimise c.x  subject to rows ``sum(coef * var) sense rhs``; variables are >= 0 (binaries <= 1)."""

    def __init__(self, name="absinthe"):
        self.name = name
        self.kind = {}            # variable -> CONTINUOUS | GENERAL | BINARY, in order of first appearance
        self.objective = {}
        self.rows = []            # (terms: dict, sense, rhs, comment or None)

    # -- building -----------------------------------------------------------------------------
    def var(self, name, kind=None):
        if name not in self.kind:
            self.kind[name] = CONTINUOUS
        if kind is not None:
            self.kind[name] = kind
        return name

    def general(self, *names):
        for name in names:
            self.var(name, GENERAL)

    def binary(self, *names):
        for name in names:
            self.var(name, BINARY)

    def minimize(self, terms):
        self.objective = self._collect(terms)

    def add(self, terms, sense, rhs, comment=None):
        assert sense in _SENSES
        self.rows.append((self._collect(terms), sense, float(rhs), comment))

    def _collect(self, terms):
        out = {}
        items = terms.items() if isinstance(terms, dict) else ((v, c) for c, v in terms)
        for name, coef in items:
            self.var(name)
            out[name] = out.get(name, 0.0) + float(coef)
        return out

    # -- LP text ------------------------------------------------------------------------------
    @staticmethod
    def _expression(terms):
        parts = []
        for name, coef in terms.items():
            sign = "-" if coef < 0 else "+"
            mag = abs(coef)
            body = name if mag == 1.0 else "%s %s" % (repr(mag), name)
            parts.append(("" if not parts and sign == "+" else sign + " ") + body)
        return " ".join(parts) if parts else "0"

    def write_lp(self, path):
        with open(path, "w") as out:
            out.write("Minimize\n" + self._expression(self.objective) + "\nSubject To\n")
            for terms, sense, rhs, comment in self.rows:
                if comment:
                    out.write("\\ " + comment + "\n")
                out.write("%s %s %s\n" % (self._expression(terms), sense, repr(rhs)))
            for title, kind in (("General", GENERAL), ("Binary", BINARY)):
                names = [n for n, k in self.kind.items() if k == kind]
                out.write(title + "\n")
                for start in range(0, len(names), 16):
                    out.write(" ".join(names[start:start + 16]) + "\n")
            out.write("End\n")

    # -- solving ------------------------------------------------------------------------------
    def solve(self, time_limit=None, gap=1e-4):
        """Returns ``Solution(objective, values, optimal)``; raises when no feasible point was found."""
        import numpy as np
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import csr_matrix
        if time_limit is None:
            time_limit = float(os.environ.get("ABSINTHE_MILP_SECONDS", "120"))
        names = list(self.kind)
        index = {name: n for n, name in enumerate(names)}
        c = np.zeros(len(names))
        for name, coef in self.objective.items():
            c[index[name]] = coef
        data, ri, ci, lo, hi = [], [], [], [], []
        for r, (terms, sense, rhs, _) in enumerate(self.rows):
            for name, coef in terms.items():
                if coef != 0.0:
                    data.append(coef)
                    ri.append(r)
                    ci.append(index[name])
            lo.append(-np.inf if sense == "<=" else rhs)
            hi.append(np.inf if sense == ">=" else rhs)
        integrality = np.array([1 if self.kind[n] != CONTINUOUS else 0 for n in names])
        upper = np.array([1.0 if self.kind[n] == BINARY else np.inf for n in names])
        matrix = csr_matrix((data, (ri, ci)), shape=(len(self.rows), len(names)))
        res = milp(c, constraints=LinearConstraint(matrix, np.array(lo), np.array(hi)), integrality=integrality,
                   bounds=Bounds(np.zeros(len(names)), upper),
                   options={"time_limit": time_limit, "mip_rel_gap": gap, "disp": False})
        if res.x is None:
            raise RuntimeError("MILP '%s': no feasible solution (%s)" % (self.name, res.message))
        values = {}
        for n, name in enumerate(names):
            values[name] = float(round(res.x[n])) if integrality[n] else float(res.x[n])
        return Solution(float(res.fun), values, res.status == 0)

    # -- comparison ---------------------------------------------------------------------------
    def canonical(self, digits=10):
        """Order-independent fingerprint material: objective, rows and variable kinds in a normal form
        (terms sorted, coefficients rounded to ``digits`` significant digits, ``<=`` rows negated to ``>=``)."""
        def num(v):
            return float("%.*e" % (digits - 1, v)) + 0.0

        def line(terms, sense, rhs):
            flip = -1.0 if sense == "<=" else 1.0
            body = tuple(sorted((n, num(flip * c)) for n, c in terms.items() if c != 0.0))
            return (body, "=" if sense == "=" else ">=", num(flip * rhs))
        rows = sorted(line(t, s, r) for t, s, r, _ in self.rows)
        used = {n for t, _, _, _ in self.rows for n, c in t.items() if c != 0.0} | set(self.objective)
        return {"objective": sorted((n, num(c)) for n, c in self.objective.items() if c != 0.0), "rows": rows,
                "general": sorted(n for n, k in self.kind.items() if k == GENERAL and n in used),
                "binary": sorted(n for n, k in self.kind.items() if k == BINARY and n in used)}


class Solution:
    def __init__(self, objective, values, optimal=True):
        self.objective, self.values, self.optimal = objective, values, optimal

    def __getitem__(self, name):
        return self.values[name]


# ---- file formats ----------------------------------------------------------------------------
_TERM = re.compile(r"([+-]?)\s*(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)?\s*([A-Za-z_%#][\w%#.]*)?")


def parse_expression(text):
    """'2 a - b + 3' -> ({var: coef}, constant)"""
    coefs, constant = {}, 0.0
    pos, text = 0, text.strip()
    while pos < len(text):
        m = _TERM.match(text, pos)
        if not m or m.end() == pos:
            if text[pos].isspace():
                pos += 1
                continue
            raise ValueError("cannot parse '%s' at %d" % (text, pos))
        sign, number, name = m.groups()
        pos = m.end()
        if number is None and name is None:
            continue
        value = (-1.0 if sign == "-" else 1.0) * (float(number) if number is not None else 1.0)
        if name is None:
            constant += value
        else:
            coefs[name] = coefs.get(name, 0.0) + value
    return coefs, constant


def read_lp(path):
    """CPLEX-LP text (sections Minimize / Subject To / General / Binary / End, ``\\`` comments) -> Model."""
    model = Model(os.path.splitext(os.path.basename(path))[0])
    section, objective = None, ""
    with open(path) as handle:
        for raw in handle:
            # a negative model parameter prints as "- -2.2e-07 x" (ref stencil_optimizer.py:790-793)
            line = raw.split("\\")[0].replace("- -", "+ ").replace("+ -", "- ").strip()
            if not line:
                continue
            key = line.lower()
            if key in ("minimize", "subject to", "general", "binary", "end", "bounds"):
                section = key
                continue
            if section == "minimize":
                objective += " " + line
            elif section == "subject to":
                if ":" in line:
                    line = line.split(":", 1)[1]
                m = re.search(r"(<=|>=|=<|=>|=)", line)
                lhs, sense, rhs = line[:m.start()], m.group(1), line[m.end():]
                terms, lk = parse_expression(lhs)
                right, rk = parse_expression(rhs)
                for name, value in right.items():
                    terms[name] = terms.get(name, 0.0) - value
                model.add(terms, {"=<": "<=", "=>": ">="}.get(sense, sense), rk - lk)
            elif section == "general":
                model.general(*line.split())
            elif section == "binary":
                model.binary(*line.split())
    if objective.strip().lower().startswith("obj:"):
        objective = objective.split(":", 1)[1]
    model.minimize(parse_expression(objective)[0])
    return model


def write_solution(path, model, solution):
    """The XML subset of a CPLEX ``.sol`` the optimiser reads back (ref stencil_optimizer.py:826-830, :899)."""
    with open(path, "w") as out:
        out.write('<?xml version = "1.0" encoding="UTF-8" standalone="yes"?>\n<CPLEXSolution version="1.2">\n')
        out.write(' <header problemName=%s objectiveValue="%.12g" solutionStatusString="%s"/>\n'
                  % (quoteattr(model.name), solution.objective, "optimal" if solution.optimal else "incumbent"))
        out.write(" <variables>\n")
        for n, name in enumerate(model.kind):
            out.write('  <variable name=%s index="%d" value="%.12g"/>\n' % (quoteattr(name), n, solution.values[name]))
        out.write(" </variables>\n</CPLEXSolution>\n")


def read_solution(path):
    """(objective string, {variable: value}) of a CPLEX ``.sol``."""
    dom = parse(path)
    values = {v.attributes["name"].value: float(v.attributes["value"].value)
              for v in dom.getElementsByTagName("variable")}
    return dom.getElementsByTagName("header")[0].attributes["objectiveValue"].value, values
