"""Stencil catalogue: the arithmetic of the kernels the backend executes.

The strings are the *data* of the hot path: each one
is a C++ statement list that defines ``res`` from field accesses of the form
``name(i+a,j+b,k+c)`` and is pasted verbatim into the oracle and into the CUDA
kernels, so operand order -- and therefore floating-point evaluation order --
is exactly the reference's.  They are built here from small generators instead
of being written out; ``tests/test_stencils.py`` pins every generated string
(whitespace-insensitively) to fingerprints taken from the reference's own
definitions:

* diffusion  -- ref diffusion.py:16-64, sequence :94-95, outputs :70
* advection  -- ref advection.py:16-54, sequence :84
* fastwaves  -- ref fastwaves.py:16-48, sequence :78
* fitcache   -- ref fitcache.py:16-59  (PT8/PT12/PT16/PT20)
* fitddr     -- ref fitddr.py:16-125   (IN{0..3}BD{0..2}, note the ``k88`` quirk)
"""


def _at(name, di=0, dj=0, dk=0):
    def idx(sym, off):
        return sym if off == 0 else "%s%+d" % (sym, off)
    return "%s(%s,%s,%s)" % (name, idx("i", di), idx("j", dj), idx("k", dk))


# ---------------------------------------------------------------------------------------
# COSMO horizontal diffusion: 5-point laplacian, flux limiter in i and j, masked update.
# ---------------------------------------------------------------------------------------
DIFFUSION_FIELDS = (("u", "uin"), ("v", "vin"), ("w", "winn"), ("pp", "ppin"))


def diffusion_stencils():
    table = {}
    for tag, src in DIFFUSION_FIELDS:
        lap, fli, flj, out = tag + "lap", tag + "fli", tag + "flj", tag + "out"
        table[lap] = ("auto res = -4.0 * %s + %s + %s + %s + %s;" % (
            _at(src), _at(src, 1), _at(src, -1), _at(src, 0, 1), _at(src, 0, -1)))
        table[fli] = ("auto fli = %s - %s; auto res = fli * (%s - %s) > 0.0 ? 0.0 : fli;" % (
            _at(lap, 1), _at(lap), _at(src, 1), _at(src)))
        table[flj] = ("auto flj = %s - %s; auto res = flj * (%s - %s) > 0.0 ? 0.0 : flj;" % (
            _at(lap, 0, 1), _at(lap), _at(src, 0, 1), _at(src)))
        table[out] = ("auto res = %s + %s * (%s - %s + %s - %s);" % (
            _at(src), _at("mask"), _at(fli, -1), _at(fli), _at(flj, 0, -1), _at(flj)))
    return table


def diffusion_sequence():
    return [tag + part for tag, _ in DIFFUSION_FIELDS for part in ("lap", "fli", "flj", "out")]


DIFFUSION_OUTPUTS = ["uout", "vout", "wout", "ppout"]
DIFFUSION_CONSTANTS = ["mask"]


# ---------------------------------------------------------------------------------------
# Horizontal advection: third-order upwind, direction chosen by the advecting velocity.
# ---------------------------------------------------------------------------------------
_UPWIND = ("-1.0/30.0", "- 1.0/4.0", "+ 1.0", "- 1.0/3.0", "- 1.0/2.0", "+ 1.0/20.0")


def _upwind(field, axis, first):
    """Six-tap polynomial starting at offset ``first`` along ``axis`` (0 = i, 1 = j)."""
    terms = []
    for n, coeff in enumerate(_UPWIND):
        off = [0, 0, 0]
        off[axis] = first + n
        terms.append("%s * %s" % (coeff, _at(field, *off)))
    return "(" + " ".join(terms) + ")"


def _advect(speed, field, axis):
    s = _at(speed)
    return "%s > 0.0 ? %s * %s : %s * %s" % (s, s, _upwind(field, axis, -3), s, _upwind(field, axis, -2))


def advection_stencils():
    return {
        "uatu": "auto res = 0.5 * (%s + %s);" % (_at("uin", -1), _at("uin", 1)),
        "uteu": "auto res = %s;" % _advect("uatu", "uin", 0),
        "vatu": "auto res = 0.25 * (%s + %s + %s + %s);" % (
            _at("vin", 1, -1), _at("vin", 1), _at("vin", 0, -1), _at("vin")),
        "utev": "auto res = %s + (%s);" % (_at("uteu"), _advect("vatu", "uin", 1)),
        "uatv": "auto res = 0.25 * (%s + %s + %s + %s);" % (
            _at("uin", -1, 1), _at("uin", -1), _at("uin", 0, 1), _at("uin")),
        "vteu": "auto res = %s;" % _advect("uatv", "vin", 0),
        "vatv": "auto res = 0.5 * (%s + %s);" % (_at("vin", 0, -1), _at("vin", 0, 1)),
        "vtev": "auto res = %s + (%s);" % (_at("vteu"), _advect("vatv", "vin", 1)),
    }


def advection_sequence():
    return ["uatu", "uteu", "vatu", "utev", "uatv", "vteu", "vatv", "vtev"]


ADVECTION_OUTPUTS = ["utev", "vtev"]
ADVECTION_CONSTANTS = []


# ---------------------------------------------------------------------------------------
# Fast waves u/v update: vertical interpolation + horizontal pressure gradient + divergence.
# ---------------------------------------------------------------------------------------
def _pgrad(di, dj):
    """Pressure gradient along the direction (di, dj) with the terrain-following correction."""
    return ("auto res = (%s - %s) + (%s + %s) * 0.5 * ((%s + %s) - (%s + %s)) / ((%s - %s) + (%s - %s));" % (
        _at("ppuv", di, dj), _at("ppuv"), _at("ppgc", di, dj), _at("ppgc"),
        _at("hhl", 0, 0, 1), _at("hhl"), _at("hhl", di, dj, 1), _at("hhl", di, dj),
        _at("hhl", 0, 0, 1), _at("hhl"), _at("hhl", di, dj, 1), _at("hhl", di, dj)))


def _velocity(src, tens, grad, di, dj):
    return "auto res = %s + 0.01 * (%s - %s * (2.0 / (%s + %s)));" % (
        _at(src), _at(tens), _at(grad), _at("rho", di, dj), _at("rho"))


def _destagger(vel, di, dj):
    return ("auto wgt = 0.5 * (%s + %s); auto res = wgt * %s + (1.0 - wgt) * %s;" % (
        _at("wgtfac"), _at("wgtfac", di, dj), _at(vel), _at(vel, 0, 0, -1)))


def _div_term(weight, dc, slope, vel, di, dj, sign):
    return "%s * ((%s - %s) * %s %s %s)" % (
        weight, _at(dc, di, dj, 1), _at(dc, di, dj), _at(slope, di, dj), sign, _at(vel, di, dj))


def fastwaves_stencils():
    return {
        "ppgk": "auto res = %s * %s + (1.0 - %s) * %s;" % (
            _at("wgtfac"), _at("ppuv"), _at("wgtfac"), _at("ppuv", 0, 0, -1)),
        "ppgc": "auto res = %s - %s;" % (_at("ppgk", 0, 0, 1), _at("ppgk")),
        "ppgu": _pgrad(1, 0),
        "ppgv": _pgrad(0, 1),
        "uout": _velocity("uin", "utens", "ppgu", 1, 0),
        "vout": _velocity("vin", "vtens", "ppgv", 0, 1),
        "udc": _destagger("uout", 1, 0),
        "vdc": _destagger("vout", 0, 1),
        "div": "auto res = %s + %s + %s + %s;" % (
            _div_term("0.1", "udc", "dzdx", "uout", 0, 0, "+"),
            _div_term("0.1", "udc", "dzdx", "uout", -1, 0, "-"),
            _div_term("0.2", "vdc", "dzdy", "vout", 0, 0, "+"),
            _div_term("0.3", "vdc", "dzdy", "vout", 0, -1, "-")),
    }


def fastwaves_sequence():
    return ["ppgk", "ppgc", "ppgu", "ppgv", "uout", "vout", "udc", "vdc", "div"]


FASTWAVES_OUTPUTS = ["uout", "vout", "div"]
FASTWAVES_CONSTANTS = ["wgtfac", "hhl", "rho", "dzdx", "dzdy"]


# ---------------------------------------------------------------------------------------
# Training stencils for the fast-memory (fitcache) and slow-memory (fitddr) model terms.
# ---------------------------------------------------------------------------------------
_FACES = ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))
_DIAG_XY = ((-1, -1, 0), (1, 1, 0), (-1, 1, 0), (1, -1, 0))
_DIAG_XZ = ((-1, 0, -1), (1, 0, 1), (-1, 0, 1), (1, 0, -1))
_DIAG_YZ = ((0, -1, -1), (0, 1, 1), (0, -1, 1), (0, 1, -1))
FITCACHE_SHAPES = {"PT8": (), "PT12": (_DIAG_XY,), "PT16": (_DIAG_XY, _DIAG_XZ),
                   "PT20": (_DIAG_XY, _DIAG_XZ, _DIAG_YZ)}


def fitcache_stencils(variant):
    """Nine chained stencils t0..t8 over the single input i0 (ref fitcache.py:16-59)."""
    diagonals = [off for ring in FITCACHE_SHAPES[variant] for off in ring]
    tail = "".join(" - " + _at("i0", *o) for o in _FACES) + "".join(" + " + _at("i0", *o) for o in diagonals)
    table = {}
    for n in range(9):
        head = _at("i0") if n == 0 else _at("t%d" % (n - 1))
        table["t%d" % n] = "auto res = " + head + tail + ";"
    return table


def fitcache_sequence():
    return ["t%d" % n for n in range(9)]


def _ddr_field(axis, n):
    return "k88" if (axis == "k" and n == 8) else "%s%d" % (axis, n)


def fitddr_stencils(variant):
    """Nine chained outputs o0..o8, IN fresh inputs per stencil, BD diagonal halo width."""
    assert variant[:2] == "IN" and variant[3:5] == "BD", variant
    n_in, width = int(variant[2]), int(variant[5])
    table = {}
    for n in range(9):
        terms = [] if n == 0 else [_at("o%d" % (n - 1))]
        for axis in "ijk"[:n_in]:
            field = _ddr_field(axis, n)
            term = _at(field)
            for d in range(1, width + 1):
                term += " - " + _at(field, -d, -d, -d) + " - " + _at(field, d, d, d)
            terms.append(term)
        table["o%d" % n] = "auto res = " + (" + ".join(terms) if terms else "0.0") + ";"
    return table


def fitddr_sequence():
    return ["o%d" % n for n in range(9)]


FITDDR_VARIANTS = ["IN0BD0", "IN1BD0", "IN2BD0", "IN3BD0", "IN1BD1", "IN2BD1", "IN3BD1",
                   "IN1BD2", "IN2BD2", "IN3BD2"]
FITCACHE_VARIANTS = ["PT8", "PT12", "PT16", "PT20"]


def fingerprint(code):
    """Whitespace-insensitive identity of a stencil string (used by the golden tests)."""
    import hashlib
    return hashlib.sha256("".join(code.split()).encode()).hexdigest()[:16]
