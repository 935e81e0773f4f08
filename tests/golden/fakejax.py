"""A tiny NumPy-backed stand-in for the JAX API surface that AutoPDEx's hot path touches, so that the
UNMODIFIED reference modules under /root/reference can be imported and EXECUTED in this container
(JAX itself is not installable here).  Test infrastructure for generating golden fixtures only --
never imported by the product or by the tests that run on the GPU box.

What is faithful: every line of reference Python (index arithmetic, masks, control flow, weak forms,
shape functions, SciPy calls) runs as written.  What is substituted: XLA (plain NumPy), `vmap`/`lax.map`
(Python loops), and the AD engine -- `jacfwd`/`jacrev` are evaluated numerically:
  * with respect to the evaluation point x of an ansatz: through the reference's own custom-JVP rule
    (spaces.py:12164-12187) by passing a tagged point, or by complex step for plain functions (exact);
  * with respect to a deformation gradient etc. (innermost, real data): complex step (exact);
  * with respect to dof arrays (outer levels): central differences, with an exactness shortcut when
    the function is linear / quadratic in the argument (weak forms are linear in the test dofs).
Accuracy of reference tangents obtained this way: ~1e-9 relative; residuals: ~1e-14.
"""
import sys
import types

import numpy as np
import scipy.linalg

REFERENCE_ROOT = "/root/reference"


# ---- arrays with the `.at[idx]` interface -----------------------------------------------------------
class JArray(np.ndarray):
    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _idx(self):
        i = self.idx
        return np.asarray(i) if isinstance(i, np.ndarray) else i

    def get(self, **kw):
        return J(np.asarray(self.arr)[self._idx()])

    def set(self, v):
        out = np.array(self.arr, copy=True)
        if np.iscomplexobj(v) and not np.iscomplexobj(out):
            out = out.astype(complex)
        out[self._idx()] = v
        return J(out)

    def add(self, v):
        out = np.array(self.arr, copy=True)
        if np.iscomplexobj(v) and not np.iscomplexobj(out):
            out = out.astype(complex)
        np.add.at(out, self._idx(), v)
        return J(out)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


def J(x):
    if isinstance(x, XPoint):
        return x
    if isinstance(x, np.ndarray):
        return x.view(JArray)
    if isinstance(x, (np.generic,)):
        return np.asarray(x).view(JArray)
    if isinstance(x, tuple):
        return tuple(J(v) for v in x)
    if isinstance(x, list):
        return [J(v) for v in x]
    return x


def _wrap_fn(fn):
    def w(*a, **k):
        return J(fn(*a, **k))
    w.__name__ = getattr(fn, "__name__", "fn")
    return w


class _NS(types.ModuleType):
    """module whose attributes fall through to a NumPy namespace, functions wrapped to return JArray"""

    def __init__(self, name, base, extra=None):
        super().__init__(name)
        self._base = base
        for k, v in (extra or {}).items():
            setattr(self, k, v)

    def __getattr__(self, name):
        v = getattr(self._base, name)
        if callable(v) and not isinstance(v, type):
            v = _wrap_fn(v)
        setattr(self, name, v)
        return v


# ---- numerical AD ---------------------------------------------------------------------------------------
class XPoint(np.ndarray):
    """evaluation point carrying a unit tangent; consumed by FakeCustomJVP (ansatz with overwrite_diff)"""
    tangent = None


class _JvpOut:
    """tangent produced by a custom-JVP rule for a tagged evaluation point; supports the LINEAR operations the
    reference applies to an ansatz before the derivative is taken (shps @ local_dofs, [0], scaling, sums)"""

    def __init__(self, tangent):
        self.tangent = tangent

    def __matmul__(self, other):
        return _JvpOut(J(np.asarray(self.tangent) @ np.asarray(other)))

    def __rmatmul__(self, other):
        return _JvpOut(J(np.asarray(other) @ np.asarray(self.tangent)))

    def __getitem__(self, idx):
        return _JvpOut(J(np.asarray(self.tangent)[idx]))

    def __mul__(self, other):
        return _JvpOut(J(np.asarray(self.tangent) * other))
    __rmul__ = __mul__

    def __add__(self, other):
        return _JvpOut(J(np.asarray(self.tangent) + (np.asarray(other.tangent) if isinstance(other, _JvpOut) else 0.0)))
    __radd__ = __add__


_CSTEP = 1e-30


class FakeCustomJVP:
    def __init__(self, fun):
        self.fun, self.rule = fun, None
        self.__name__ = getattr(fun, "__name__", "custom_jvp")

    def defjvp(self, rule):
        self.rule = rule
        return rule

    def __call__(self, *args):
        tagged = [i for i, a in enumerate(args) if isinstance(a, XPoint) and a.tangent is not None]
        if tagged:
            prim = tuple(np.asarray(a).view(JArray) if isinstance(a, np.ndarray) else a for a in args)
            tang = tuple(J(np.asarray(args[i].tangent)) if i in tagged else
                         (J(np.zeros_like(np.asarray(a, dtype=float))) if (i == 0 or isinstance(a, np.ndarray)) else None)
                         for i, a in enumerate(args))
            tang = tuple(t if t is not None else None for t in tang)
            _, t_out = self.rule(prim, tang)
            return _JvpOut(t_out)
        cplx = [i for i, a in enumerate(args) if isinstance(a, np.ndarray) and np.iscomplexobj(a)
                and np.any(np.imag(a) != 0)]
        if cplx:
            prim = tuple(J(np.real(a)) if isinstance(a, np.ndarray) else a for a in args)
            tang = tuple(J(np.imag(a) / _CSTEP) if (isinstance(a, np.ndarray) and np.iscomplexobj(a))
                         else (J(np.zeros_like(np.asarray(a, dtype=float))) if isinstance(a, np.ndarray) else None)
                         for a in args)
            p_out, t_out = self.rule(prim, tang)
            return J(np.asarray(p_out) + 1j * _CSTEP * np.asarray(t_out))
        return self.fun(*args)


def _flatten(x):
    if isinstance(x, dict):
        keys = list(x.keys())
        parts = [np.asarray(x[k], dtype=float) for k in keys]
        sizes = [p.size for p in parts]

        def unflat(v):
            out, o = {}, 0
            for k, p, s in zip(keys, parts, sizes):
                out[k] = J(v[o:o + s].reshape(p.shape))
                o += s
            return out
        return np.concatenate([p.ravel() for p in parts]), unflat
    a = np.asarray(x)
    return a.ravel().astype(complex if np.iscomplexobj(a) else float), lambda v: J(v.reshape(a.shape))


def _out_flat(y):
    if isinstance(y, dict):
        return np.concatenate([np.asarray(y[k]).ravel() for k in y.keys()]), ("dict", {k: np.shape(y[k]) for k in y})
    return np.asarray(y).ravel(), ("arr", np.shape(y))


def _assemble_jac(cols, out_info, x):
    """cols[j] = d out / d x_j (flat).  Result structure: out.shape + x.shape (dicts of dicts for pytrees)."""
    Jm = np.stack(cols, axis=-1)                                   # (n_out, n_in)
    kind, oshape = out_info

    def split_in(mat_rows, lead_shape):
        if isinstance(x, dict):
            out, o = {}, 0
            for k in x.keys():
                s = np.asarray(x[k]).size
                out[k] = J(mat_rows[..., o:o + s].reshape(lead_shape + np.shape(x[k])))
                o += s
            return out
        return J(mat_rows.reshape(lead_shape + np.shape(x)))
    if kind == "dict":
        out, o = {}, 0
        for k, shp in oshape.items():
            s = int(np.prod(shp)) if shp else 1
            out[k] = split_in(Jm[o:o + s], tuple(shp))
            o += s
        return out
    return split_in(Jm, tuple(oshape))


_AD_DEPTH = [0]
POINT_MODE = "complex"   # "fd": differentiate plain functions of the evaluation point by Richardson differences


def jac(fun, argnums=0, has_aux=False, **_):
    def dfun(*args, **kwargs):
        x = args[argnums]
        call = lambda v: fun(*args[:argnums], v, *args[argnums + 1:], **kwargs)
        xa = np.asarray(x) if not isinstance(x, dict) else None
        # (1) derivative with respect to a small evaluation point: custom-JVP tag, else complex step
        if xa is not None and xa.ndim <= 1 and xa.size <= 4 and not np.iscomplexobj(xa):
            cols, info, ok = [], None, True
            for k in range(max(xa.size, 1)):
                p = np.array(xa, dtype=float).view(XPoint)
                t = np.zeros(xa.shape)
                t.reshape(-1)[k] = 1.0
                p.tangent = t
                y = call(p)
                if not isinstance(y, _JvpOut):
                    ok = False
                    break
                f, info = _out_flat(y.tangent)
                cols.append(f)
            if ok:
                return _assemble_jac(cols, info, xa)
        flat, unflat = _flatten(x)
        n = flat.size
        # (2) plain functions of a small evaluation point, and innermost real data (strain energies): complex step
        small_point = xa is not None and xa.ndim <= 1 and xa.size <= 4
        if not np.iscomplexobj(flat) and n <= 16 and ((small_point and POINT_MODE == "complex") or _is_leaf_level(fun)):
            cols, info = [], None
            for k in range(n):
                v = flat.astype(complex)
                v[k] += 1j * _CSTEP
                f, info = _out_flat(call(unflat(v)))
                cols.append(np.imag(f) / _CSTEP)
            return _assemble_jac(cols, info, x)
        # (3) dof arrays: central differences; exact shortcut for functions linear/quadratic in the argument
        _AD_DEPTH[0] += 1
        try:
            cols, info = [], None
            f0, info = _out_flat(call(unflat(flat.copy())))
            for k in range(n):
                def d(h):
                    vp, vm = flat.copy(), flat.copy()
                    vp[k] += h
                    vm[k] -= h
                    fp, _ = _out_flat(call(unflat(vp)))
                    fm, _ = _out_flat(call(unflat(vm)))
                    return (fp - fm) / (2 * h), fp + fm - 2 * f0
                d1, curv1 = d(1.0)
                d2, curv2 = d(0.5)
                scale = max(np.abs(d1).max(), np.abs(d2).max(), 1e-300)
                if np.abs(d1 - d2).max() <= 1e-11 * scale:
                    cols.append(d1)                                # linear or quadratic in x_k: exact
                else:
                    h = 1e-5 * max(1.0, abs(flat[k]))
                    # Richardson-extrapolated central difference, error O(h^4)
                    a, _ = d(h)
                    b, _ = d(h / 2)
                    cols.append((4 * b - a) / 3)
            return _assemble_jac(cols, info, x)
        finally:
            _AD_DEPTH[0] -= 1
    return dfun


def _is_leaf_level(fun):
    """Strain-energy style functions (plain functions of a small matrix) are differentiated by complex step."""
    name = getattr(fun, "__name__", "")
    return name in ("neo_hooke", "isochoric_neo_hooke", "linear_elastic_strain_energy", "<lambda>_leaf") or \
        getattr(fun, "_complex_step_ok", False)


def vmap(fun, in_axes=0, out_axes=0, **_):
    def run(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaf = next(iter(a.values())) if isinstance(a, dict) else a
                n = np.shape(leaf)[0]
                break
        outs = []
        for i in range(n):
            sl = []
            for a, ax in zip(args, axes):
                if ax is None:
                    sl.append(a)
                elif isinstance(a, dict):
                    sl.append({k: J(np.asarray(v)[i]) for k, v in a.items()})
                else:
                    sl.append(J(np.asarray(a)[i]))
            outs.append(fun(*sl))
        return _stack(outs)
    return run


def _stack(outs):
    o0 = outs[0]
    if isinstance(o0, _JvpOut):
        return _JvpOut(J(np.stack([np.asarray(o.tangent) for o in outs])))
    if isinstance(o0, tuple):
        return tuple(_stack([o[i] for o in outs]) for i in range(len(o0)))
    if isinstance(o0, dict):
        return {k: _stack([o[k] for o in outs]) for k in o0}
    return J(np.stack([np.asarray(o) for o in outs]))


def tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: tree_map(f, *[t[k] for t in trees]) for k in t0}
    if isinstance(t0, (tuple, list)):
        return type(t0)(tree_map(f, *[t[i] for t in trees]) for i in range(len(t0)))
    return f(*trees)


class BCOO:
    def __init__(self, args, shape=None, **_):
        self.data, self.indices = J(np.asarray(args[0])), J(np.asarray(args[1]))
        self.shape = tuple(shape)

    @property
    def nse(self):
        return self.data.shape[0]

    def __add__(self, other):
        return BCOO((np.concatenate([self.data, other.data]), np.concatenate([self.indices, other.indices])), shape=self.shape)
    __radd__ = __add__


def _sparse_empty(shape, dtype=float, index_dtype=np.int64, **_):
    return BCOO((np.zeros(0, dtype=dtype), np.zeros((0, 2), dtype=index_dtype)), shape=shape)


class FrozenDict(dict):
    def __hash__(self):
        return id(self)


def _while_loop(cond, body, carry):
    while cond(carry):
        carry = body(carry)
    return carry


def _fori_loop(lo, hi, body, carry):
    for i in range(int(lo), int(hi)):
        carry = body(i, carry)
    return carry


def _cond(pred, tf, ff, *ops, **kw):
    if "operand" in kw:                       # the older keyword form jax.lax.cond(pred, tf, ff, operand=x)
        ops = (kw["operand"],)
    return tf(*ops) if bool(pred) else ff(*ops)


def _jvp(fun, primals, tangents):
    """jax.jvp for one array argument by a central difference along the tangent (dae.newton_solver builds its dense
    Jacobian column by column this way; exact to rounding for the linear test systems it is used on here)."""
    (x,), (v,) = primals, tangents
    x, v = np.asarray(x, dtype=float), np.asarray(v, dtype=float)
    h = 1e-6 * max(1.0, float(np.abs(x).max()))
    fp, fm = np.asarray(fun(J(x + h * v))), np.asarray(fun(J(x - h * v)))
    return J(np.asarray(fun(J(x)))), J((fp - fm) / (2 * h))


def _custom_root(f, initial_guess, solve, tangent_solve, has_aux=False):
    """jax.lax.custom_root: the forward pass is solve(f, initial_guess); the implicit-differentiation rule is not needed."""
    return solve(f, initial_guess)


def _lax_map(f, xs, batch_size=None):
    return _stack([f(J(np.asarray(xs)[i])) for i in range(np.shape(xs)[0])])


class _Dummy:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]                      # used as a decorator
        raise NotImplementedError("fakejax: %s is not emulated" % self._name)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy(self._name + "." + name)


def install():
    """Register the stand-in modules and make `import autopdex` resolve to the reference tree."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "_is_fake", False):
        raise RuntimeError("a real jax is importable; the fake must not shadow it")
    jnp = _NS("jax.numpy", np, {
        "ndarray": np.ndarray, "int_": np.int64, "float64": np.float64, "bool_": np.bool_, "float_": np.float64,
        "array": _wrap_fn(np.array), "asarray": _wrap_fn(lambda x, dtype=None: _asarray(x, dtype)),
        "linalg": _NS("jax.numpy.linalg", np.linalg),
        "invert": _wrap_fn(np.invert),
        "astype": _wrap_fn(lambda x, dtype, **k: np.asarray(x).astype(dtype)),     # jnp.astype accepts Python scalars
        "einsum": _wrap_fn(lambda sub, *ops, **k: np.einsum(sub.replace(" ", ""), *ops, **k)), "pi": np.pi, "inf": np.inf, "nan": np.nan, "newaxis": None,
    })
    class _Mod(types.ModuleType):
        def __getattr__(self, name):          # anything the hot path never calls
            if name.startswith("__"):
                raise AttributeError(name)
            return _Dummy(self.__name__ + "." + name)
    jax = _Mod("jax")
    jax._is_fake = True
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.vmap = vmap
    jax.jacfwd = jac
    jax.jacrev = jac
    jax.grad = jac
    jax.hessian = lambda f, **kw: jac(jac(f, **kw), **kw)
    jax.jvp = _jvp
    jax.vjp = None
    jax.linearize = None
    jax.custom_jvp = FakeCustomJVP
    jax.pure_callback = lambda f, shape, *a, **k: J(np.asarray(f(*a)))
    jax.ShapeDtypeStruct = lambda shape=None, dtype=None: (shape, dtype)
    jax.eval_shape = lambda f, *a: types.SimpleNamespace(shape=np.shape(f(*a)))
    jax.ensure_compile_time_eval = None
    jax.Array = np.ndarray
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.debug = types.SimpleNamespace(print=lambda *a, **k: None)
    lax = types.ModuleType("jax.lax")
    lax.while_loop, lax.fori_loop, lax.cond, lax.map = _while_loop, _fori_loop, _cond, _lax_map
    lax.stop_gradient = lambda x: x
    lax.custom_root = _custom_root
    jax.lax = lax
    tree = types.ModuleType("jax.tree")
    tree.map = tree_map
    jax.tree = tree
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.tree_map = tree_map
    # autopdex.dae registers its state dataclasses / the manager as pytrees at import time: bookkeeping of the tracer only
    tree_util.register_dataclass = lambda cls=None, **kw: (cls if cls is not None else (lambda c: c))
    tree_util.register_pytree_node = lambda *a, **k: None
    jax.tree_util = tree_util
    exp = types.ModuleType("jax.experimental")
    sparse = types.ModuleType("jax.experimental.sparse")
    sparse.BCOO, sparse.empty = BCOO, _sparse_empty
    exp.sparse = sparse
    jax.experimental = exp
    jscipy = types.ModuleType("jax.scipy")
    jscipy.linalg = _NS("jax.scipy.linalg", scipy.linalg)
    jscipy.sparse = types.SimpleNamespace(linalg=types.SimpleNamespace())
    jax.scipy = jscipy
    jax.random = _Mod("jax.random")
    jax.random.key = lambda seed: seed
    jax.random.PRNGKey = lambda seed: seed
    jax.random.split = lambda key, num=2: tuple(range(num))
    flax = types.ModuleType("flax")
    core = types.ModuleType("flax.core")
    core.FrozenDict = FrozenDict
    flax.core = core
    jaxopt = _Mod("jaxopt")
    mods = {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.tree": tree, "jax.tree_util": tree_util,
            "jax.experimental": exp, "jax.experimental.sparse": sparse, "jax.scipy": jscipy,
            "jax.scipy.linalg": jscipy.linalg, "jax.random": jax.random, "flax": flax, "flax.core": core,
            "jaxopt": jaxopt}
    sys.modules.update(mods)
    for name in ("vmap", "jacrev", "jacfwd", "hessian", "jvp", "linearize", "vjp", "custom_jvp", "lax", "random"):
        pass
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return jax


def _asarray(x, dtype=None):
    if isinstance(x, (list, tuple)) and any(isinstance(v, np.ndarray) and np.iscomplexobj(v) for v in _leaves(x)):
        return np.asarray(x, dtype=complex)
    return np.asarray(x, dtype=dtype)


def _leaves(x):
    for v in x:
        if isinstance(v, (list, tuple)):
            yield from _leaves(v)
        else:
            yield v
