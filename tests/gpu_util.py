"""Helpers for the GPU parity tests: oracle set dicts -> backend SetSpecs."""
import numpy as np

from autopdex_b200 import backend

_FAMILY = {"line2": "quad_brick", "line3": "quad_brick", "quad4": "quad_brick", "quad9": "quad_brick",
           "hex8": "quad_brick", "hex27": "quad_brick", "tri3": "tri_tet", "tri6": "tri_tet",
           "tet4": "tri_tet", "tet10": "tri_tet"}


def to_setspec(st):
    m = st["model"]
    params = {k: v for k, v in m.items() if k not in ("name", "mode") and not k.startswith("_") and v is not None}
    if st["kind"] == "intpoint":
        return backend.SetSpec("intpoint", m["name"], st["conn"], mode=m.get("mode"), params=params,
                               tables=(st["N"], st["dNdx"], st["w"]))
    return backend.SetSpec(st["kind"], m["name"], st["conn"], family=_FAMILY[st["etype"]], gp=st["gp"],
                           mode=m.get("mode"), params=params)


def make_plan(p, settings=None):
    sets = [to_setspec(s) for s in p["sets"]]
    coords = np.asarray(p["coords"], dtype=np.float64)
    plan = backend.Plan(coords.shape[1], coords.shape[0], p["nf"], sets, p["mask"])
    plan.set_coords(coords)
    settings = settings or p.get("settings") or {}
    if "time increment" in settings:
        plan.set_time_increment(settings["time increment"])
    if "dofs n" in settings:
        plan.set_dofs_n(settings["dofs n"])
    return plan
