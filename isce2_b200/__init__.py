"""isce2_b200 -- B200-native (sm_100a, hand-written FP64 CUDA) zero-Doppler radar geometry for ISCE2:
topozero (rdr2geo), geo2rdr and geozero (geocode) behind the reference's Component API.

    from isce2_b200 import createTopozero, createGeo2rdr       # same objects as zerodop.topozero / zerodop.geo2rdr

or, to switch an existing ISCE2 workflow (topsApp / stripmapApp / topsStack runTopo, runGeo2rdr) without touching it:

    import isce2_b200; isce2_b200.install_as_zerodop()          # before the workflow imports zerodop.*

The compute path is libb200geom.so (isce2_b200/csrc, C ABI in include/b200geom.h).  There is no CPU fallback.
"""
import sys
import types

__version__ = "0.1.0"


def createTopozero():
    from .topozero import createTopozero as f
    return f()


def createGeo2rdr(name=''):
    from .geo2rdr import createGeo2rdr as f
    return f(name)


def createGeozero():
    from .geozero import createGeozero as f
    return f()


def createResamp_slc():
    from .resamp_slc import createResamp_slc as f
    return f()


def install_as_zerodop():
    """Make ``from zerodop.topozero import createTopozero`` / ``from zerodop.geo2rdr import createGeo2rdr`` resolve to
    this package (components/zerodop/topozero/__init__.py:33-35, components/zerodop/geo2rdr/__init__.py:3-5)."""
    from . import geo2rdr as g, geozero as z, topozero as t
    pkg = sys.modules.get("zerodop")
    if pkg is None:
        pkg = types.ModuleType("zerodop")
        pkg.__path__ = []
        sys.modules["zerodop"] = pkg
    for name, mod, factory in (("topozero", t, "createTopozero"), ("geo2rdr", g, "createGeo2rdr"),
                               ("geozero", z, "createGeozero")):
        # zerodop.<name> is a package in the reference (its class lives in zerodop/<name>/<Name>.py): both import forms,
        # `from zerodop.topozero import createTopozero` and `from zerodop.topozero.Topozero import Topo`, must resolve
        m = types.ModuleType(f"zerodop.{name}")
        m.__path__ = []
        setattr(m, factory, getattr(mod, factory))
        sub = mod.__name__.rsplit(".", 1)[-1].capitalize()  # Topozero / Geo2rdr / Geozero
        setattr(m, sub, mod)
        m.__dict__.update({k: v for k, v in mod.__dict__.items() if k in ("Topo", "Geocode")})  # (Geo2rdr names the submodule)
        sys.modules[f"zerodop.{name}"] = m
        sys.modules[f"zerodop.{name}.{sub}"] = mod
        setattr(pkg, name, m)
    return pkg
