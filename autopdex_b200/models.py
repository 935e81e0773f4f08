"""Model objects of the b200 backend (host side).

Same factory names and argument meaning as autopdex/models.py for the built-in models of the
hot path; instead of JAX closures they return small *descriptors* that the backend maps to the
closed-form CUDA element kernels:

  poisson_weak                         models.py:96-134
  linear_elasticity_weak               models.py:510-635
  neumann_weak                         models.py:744-779
  hyperelastic_steady_state_weak       models.py:917-1000   (strain energy: neo_hooke, :1122-1146)
  forward_backward_euler_weak          models.py:1946-2010
  mixed_reference_domain_potential     models.py:1188-1269  (tagged integrands only)
  isoparametric_domain_element_galerkin / isoparametric_surface_element_galerkin   models.py:1616-1850

`recognise(model)` also accepts the *reference's own closures* (when AutoPDEx/JAX is importable):
they are identified by __qualname__ and their closure cells, never called.  Anything else --
in particular a user-written integrand -- is rejected with ValueError when static_settings are
read; there is no CPU path to fall back to.
"""
import numpy as np

from . import spaces


class WeakForm:
    """name in {'poisson_weak','poisson_potential','linear_elasticity','neo_hooke','neumann','capacity'};
    funs maps backend parameter names to Python callables f(x) / f(x, settings) or constants."""

    def __init__(self, name, funs, mode=None):
        self.name, self.funs, self.mode = name, dict(funs), mode

    def __repr__(self):
        return "<b200 weak form %s%s>" % (self.name, "" if self.mode is None else " (%s)" % self.mode)


class ElementModel:
    """An isoparametric element: kind 'domain' | 'surface', weak form, shape family, Gauss rule.
    physical_x: coefficient callables receive the PHYSICAL Gauss-point coordinate (mixed potential
    route, models.py:1244-1254) instead of the reference coordinate (user element route, :1671-1679)."""

    def __init__(self, kind, weak, family, gp, physical_x=False, field=None):
        self.kind, self.weak, self.family, self.physical_x, self.field = kind, weak, family, physical_x, field
        xi, w = gp
        self.gp = (np.asarray(xi, dtype=np.float64), np.asarray(w, dtype=np.float64))


def _unsupported(what):
    raise ValueError("b200 backend: %s is not supported (supported: Poisson/heat, linear elasticity, neo-Hooke, "
                     "Neumann loads, backward-Euler capacity term on Q1/Q2 line/quad/hex and P1/P2 tri/tet elements); "
                     "there is no CPU fallback" % what)


# ---- weak forms ------------------------------------------------------------------------------
def poisson_weak(coefficient_fun=lambda x, settings: 1.0, source_fun=None):
    return WeakForm("poisson_weak", {"coefficient": coefficient_fun, "source": source_fun})


def linear_elasticity_weak(youngs_mod_fun, poisson_ratio_fun, mode, volume_load_fun=None):
    if mode not in ("plain strain", "plain stress", "3d"):
        raise AssertionError("'mode' for linear elasticity not properly set.")
    return WeakForm("linear_elasticity", {"youngs_modulus": youngs_mod_fun, "poisson_ratio": poisson_ratio_fun,
                                          "body_load": volume_load_fun}, mode)


def neumann_weak(neumann_fun):
    return WeakForm("neumann", {"traction": neumann_fun})


def neo_hooke(F, param):
    """Ciarlet-type neo-Hookean energy psi = mu/2 (tr C - 3 - 2 ln J) + lam/4 (J^2 - 1 - 2 ln J)
    (models.py:1122-1146).  Token: the backend uses its closed-form P and dP/dF."""
    raise RuntimeError("neo_hooke is evaluated on the device by the b200 backend")


def linear_elastic_strain_energy(F, param):
    """Small-strain energy psi = lam/2 tr(eps)^2 + mu eps:eps, eps = sym(F - 1) (models.py:1167-1185).  Token: inside
    hyperelastic_steady_state_weak its first Piola-Kirchhoff stress is sigma = lam tr(eps) 1 + 2 mu eps, i.e. the linear
    elasticity kernel with the isotropic tensor in the mesh's dimension ('plain strain': eps_33 = 0)."""
    raise RuntimeError("linear_elastic_strain_energy is evaluated on the device by the b200 backend")


def hyperelastic_steady_state_weak(strain_energy_fun, youngs_mod_fun, poisson_ratio_fun, mode, volume_load_fun=None):
    energy = getattr(strain_energy_fun, "__name__", None)
    if energy not in ("neo_hooke", "linear_elastic_strain_energy"):
        _unsupported("strain energy %r (only models.neo_hooke and models.linear_elastic_strain_energy)" % (strain_energy_fun,))
    if mode not in ("plain strain", "3d"):
        raise AssertionError("Hyperelastic model supports only 'plain strain' and '3d' modes.")
    if energy == "linear_elastic_strain_energy":
        # device mode 'lame' (APDX_MODE_LAME) = the isotropic tensor lam 1x1 + 2 mu I_sym in the mesh's dimension; on a 2-D
        # mesh that is the customary plane-strain matrix -- not the one of linear_elasticity_weak 'plain strain', whose
        # shear entry the reference doubles (models.py:570-577)
        return WeakForm("linear_elasticity", {"youngs_modulus": youngs_mod_fun, "poisson_ratio": poisson_ratio_fun,
                                              "body_load": volume_load_fun}, "lame")
    return WeakForm("neo_hooke", {"youngs_modulus": youngs_mod_fun, "poisson_ratio": poisson_ratio_fun,
                                  "body_load": volume_load_fun}, mode)


def forward_backward_euler_weak(inertia_coeff_fun):
    return WeakForm("capacity", {"coefficient": inertia_coeff_fun})


# ---- tagged integrands for the mixed-potential route ---------------------------------------------
class poisson_potential:
    """Tagged integrand  Pi = 1/2 c grad(phi).grad(phi) - f phi  for mixed_reference_domain_potential.

    The README integrand (examples/miscellaneous/short_example.py:23-33) is of this form.  A
    user-written integrand cannot be inspected, so the b200 backend only accepts tagged ones.
    source_fun / coefficient_fun receive physical coordinates; with vectorized=True they are
    called once with an (n_points, dim) array.  The object is also a valid JAX integrand (same
    signature as the reference expects) when JAX is installed, so the same settings run on the
    reference 'scipy' backend for parity checks.
    """

    def __init__(self, field, source_fun=None, coefficient_fun=None, vectorized=True):
        self.field, self.source_fun, self.coefficient_fun, self.vectorized = field, source_fun, coefficient_fun, vectorized

    def __call__(self, x_int, ansatz_fun, settings, static_settings, elem_number, set):
        import jax  # only reachable on the reference backend
        x = ansatz_fun["physical coor"](x_int)
        phi_fun = ansatz_fun[self.field]
        dphi = jax.jacrev(phi_fun)(x_int)
        c = 1.0 if self.coefficient_fun is None else self.coefficient_fun(x)
        f = 0.0 if self.source_fun is None else self.source_fun(x)
        return 0.5 * c * dphi @ dphi - f * phi_fun(x_int)


class poisson_residual:
    """Tagged weak-form integrand  c grad(phi).grad(dphi) - f dphi  for mixed_reference_domain_residual
    ('user residual' assembling mode, models.py:1272-1357 / assembler.py:1217-1308): the variation of poisson_potential,
    same element arithmetic.  Also a valid JAX integrand for the reference (signature of models.py:1330-1338)."""

    def __init__(self, field, source_fun=None, coefficient_fun=None, vectorized=True):
        self.field, self.source_fun, self.coefficient_fun, self.vectorized = field, source_fun, coefficient_fun, vectorized

    def __call__(self, x_int, trial_ansatz, test_ansatz, settings, static_settings, elem_number, set):
        import jax  # only reachable on the reference backend
        x = trial_ansatz["physical coor"](x_int)
        dphi = jax.jacfwd(trial_ansatz[self.field])(x_int)
        dtest = jax.jacfwd(test_ansatz[self.field])(x_int)
        c = 1.0 if self.coefficient_fun is None else self.coefficient_fun(x)
        f = 0.0 if self.source_fun is None else self.source_fun(x)
        return c * dphi @ dtest - f * test_ansatz[self.field](x_int)


class heat_conduction_time:
    """Tagged TIME-DEPENDENT weak-form integrand  c theta_t dtheta + k grad(theta).grad(dtheta) - f dtheta  for
    mixed_reference_domain_residual_time (models.py:1854-1914), the 'user residual' that dae.TimeSteppingManager
    assembles (dae.py:1809-1876): theta_t comes from the field's time integrator.  Coefficient callables receive the
    physical coordinate (and settings).  Also a valid JAX integrand for the reference."""

    def __init__(self, field, conductivity_fun=None, capacity_fun=None, source_fun=None, vectorized=True):
        self.field, self.conductivity_fun, self.capacity_fun, self.source_fun = field, conductivity_fun, capacity_fun, source_fun
        self.vectorized = vectorized

    def __call__(self, x_int, trial_ansatz, test_ansatz, settings, static_settings, elem_number, set):
        import jax  # only reachable on the reference backend
        t = settings["current time"]
        x = trial_ansatz["physical coor"](x_int)
        theta = trial_ansatz[self.field]
        theta_t = jax.jacfwd(lambda tt: theta(x_int, tt))(t)
        dtheta = jax.jacfwd(lambda xx: theta(xx, t))(x_int)
        test = test_ansatz[self.field]
        k = 1.0 if self.conductivity_fun is None else self.conductivity_fun(x)
        c = 1.0 if self.capacity_fun is None else self.capacity_fun(x)
        f = 0.0 if self.source_fun is None else self.source_fun(x)
        return c * theta_t * test(x_int) + k * dtheta @ jax.jacfwd(test)(x_int) - f * test(x_int)


class TimeElementModel:
    """A transient 'user residual' domain: one reference domain served by TWO device sets on the same connectivity --
    the steady part (conduction + source, positive sign: the poisson_potential kernel) and the capacity kernel, whose
    backward-difference form  c (a q + b)  is fed by the time integrator (dae.py of this package)."""

    def __init__(self, steady, capacity, field):
        self.steady, self.capacity, self.field = steady, capacity, field
        self.kind = "domain"


# ---- element factories ----------------------------------------------------------------------------
def _family(ansatz_fun):
    name = getattr(ansatz_fun, "__name__", "")
    if name == "fem_iso_line_quad_brick":
        return "quad_brick"
    if name == "fem_iso_line_tri_tet":
        return "tri_tet"
    _unsupported("ansatz function %r" % (ansatz_fun,))


def isoparametric_domain_element_galerkin(weak_form_fun, ansatz_fun, ref_int_coor, ref_int_weights,
                                          initial_config=True):
    if not initial_config:
        _unsupported("initial_config=False (updated-Lagrangian mapping)")
    weak = recognise_weak_form(weak_form_fun)
    if weak.name in ("neumann", "poisson_potential"):
        _unsupported("weak form %s inside a domain user element" % weak.name)
    # 'capacity' (forward_backward_euler_weak) inside an isoparametric element is an extension of the b200 backend:
    # the reference evaluates it through solution_structure, i.e. in 'sparse' mode only (models.py:1996-1998)
    return ElementModel("domain", weak, _family(ansatz_fun), (ref_int_coor, ref_int_weights))


def isoparametric_surface_element_galerkin(weak_form_fun, ansatz_fun, ref_int_coor, ref_int_weights,
                                           tangent_contributions, initial_config=True):
    if not initial_config:
        _unsupported("initial_config=False (follower loads)")
    weak = recognise_weak_form(weak_form_fun)
    if weak.name != "neumann":
        _unsupported("weak form %s on a surface element (only neumann_weak)" % weak.name)
    # tangent_contributions True or False: d(residual)/d(trial dofs) of neumann_weak is zero either way
    return ElementModel("surface", weak, _family(ansatz_fun), (ref_int_coor, ref_int_weights))


def mixed_reference_domain_potential(integrand_fun, ansatz_fun, ref_int_coor, ref_int_weights, mapping_key):
    if not isinstance(integrand_fun, poisson_potential):
        _unsupported("a user-written integrand (use a tagged integrand such as models.poisson_potential)")
    # multi-field dict dofs: ansatz_fun names every field (models.py:1236-1243 builds an ansatz per key); the tagged
    # integrand acts on ONE of them, and that field also carries the isoparametric mapping
    if integrand_fun.field not in ansatz_fun or mapping_key != integrand_fun.field:
        _unsupported("a potential whose field %r is not in ansatz_fun / is not the mapping key" % (integrand_fun.field,))
    weak = WeakForm("poisson_potential", {"coefficient": integrand_fun.coefficient_fun,
                                          "source": integrand_fun.source_fun})
    weak.vectorized = integrand_fun.vectorized
    return ElementModel("domain", weak, _family(ansatz_fun[integrand_fun.field]), (ref_int_coor, ref_int_weights),
                        physical_x=True, field=integrand_fun.field)


def mixed_reference_domain_residual(integrand_fun, ansatz_fun, ref_int_coor, ref_int_weights, mapping_key):
    """'user residual' route (models.py:1272-1357): tagged integrands only."""
    if not isinstance(integrand_fun, poisson_residual):
        _unsupported("a user-written weak-form integrand (use a tagged integrand such as models.poisson_residual)")
    if integrand_fun.field not in ansatz_fun or mapping_key != integrand_fun.field:
        _unsupported("a residual whose field %r is not in ansatz_fun / is not the mapping key" % (integrand_fun.field,))
    weak = WeakForm("poisson_potential", {"coefficient": integrand_fun.coefficient_fun, "source": integrand_fun.source_fun})
    weak.vectorized = integrand_fun.vectorized
    m = ElementModel("domain", weak, _family(ansatz_fun[integrand_fun.field]), (ref_int_coor, ref_int_weights),
                     physical_x=True, field=integrand_fun.field)
    m.route = "user residual"
    return m


def mixed_reference_domain_residual_time(integrand_fun, ansatz_fun, ref_int_coor, ref_int_weights, mapping_key):
    """Time-dependent 'user residual' of dae.TimeSteppingManager (models.py:1854-1914): tagged integrands only."""
    if not isinstance(integrand_fun, heat_conduction_time):
        _unsupported("a user-written time-dependent integrand (use a tagged one such as models.heat_conduction_time)")
    if list(ansatz_fun.keys()) != [integrand_fun.field] or mapping_key != integrand_fun.field:
        _unsupported("multi-field residuals")
    fam = _family(ansatz_fun[integrand_fun.field])
    steady = WeakForm("poisson_potential", {"coefficient": integrand_fun.conductivity_fun, "source": integrand_fun.source_fun})
    cap = WeakForm("capacity", {"coefficient": integrand_fun.capacity_fun if integrand_fun.capacity_fun is not None else 1.0})
    steady.vectorized = cap.vectorized = integrand_fun.vectorized
    gp = (ref_int_coor, ref_int_weights)
    return TimeElementModel(ElementModel("domain", steady, fam, gp, physical_x=True, field=integrand_fun.field),
                            ElementModel("domain", cap, fam, gp, physical_x=True, field=integrand_fun.field), integrand_fun.field)


# ---- recognition of the reference's own closures ---------------------------------------------------
def _cells(fn):
    code, clo = getattr(fn, "__code__", None), getattr(fn, "__closure__", None)
    if code is None or clo is None:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, clo):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def _made_by(qualname, factory):
    """True if `qualname` is that of a closure created directly inside `factory`
    (e.g. 'poisson_weak.<locals>.pde_fun', possibly with an enclosing prefix)."""
    return qualname.split(".")[-3:-1] == [factory, "<locals>"]


def recognise_weak_form(fn):
    if isinstance(fn, WeakForm):
        return fn
    qn = getattr(fn, "__qualname__", "")
    c = _cells(fn)
    if _made_by(qn, "poisson_weak"):
        return poisson_weak(c.get("coefficient_fun"), c.get("source_fun"))
    if _made_by(qn, "linear_elasticity_weak"):
        return linear_elasticity_weak(c.get("youngs_mod_fun"), c.get("poisson_ratio_fun"), c.get("mode"),
                                      c.get("volume_load_fun"))
    if _made_by(qn, "hyperelastic_steady_state_weak"):
        return hyperelastic_steady_state_weak(c.get("strain_energy_fun"), c.get("youngs_mod_fun"),
                                              c.get("poisson_ratio_fun"), c.get("mode"), c.get("volume_load_fun"))
    if _made_by(qn, "neumann_weak"):
        return neumann_weak(c.get("neumann_fun"))
    if _made_by(qn, "forward_backward_euler_weak"):
        return forward_backward_euler_weak(c.get("inertia_coeff_fun"))
    _unsupported("model %r" % (fn,))


def recognise(model):
    """ElementModel / WeakForm for one entry of static_settings['model'], or ValueError."""
    if isinstance(model, (ElementModel, WeakForm, TimeElementModel)):
        return model
    qn = getattr(model, "__qualname__", "")
    c = _cells(model)
    if _made_by(qn, "isoparametric_domain_element_galerkin"):
        return isoparametric_domain_element_galerkin(c.get("weak_form_fun"), c.get("ansatz_fun"), c.get("ref_int_coor"),
                                                     c.get("ref_int_weights"), c.get("initial_config", True))
    if _made_by(qn, "isoparametric_surface_element_galerkin"):
        return isoparametric_surface_element_galerkin(c.get("weak_form_fun"), c.get("ansatz_fun"), c.get("ref_int_coor"),
                                                      c.get("ref_int_weights"), c.get("tangent_contributions", False),
                                                      c.get("initial_config", True))
    if _made_by(qn, "mixed_reference_domain_potential"):
        return mixed_reference_domain_potential(c.get("integrand_fun"), c.get("ansatz_fun"), c.get("ref_int_coor"),
                                                c.get("ref_int_weights"), c.get("mapping_key"))
    if _made_by(qn, "mixed_reference_domain_residual"):
        return mixed_reference_domain_residual(c.get("integrand_fun"), c.get("ansatz_fun"), c.get("ref_int_coor"),
                                               c.get("ref_int_weights"), c.get("mapping_key"))
    if _made_by(qn, "mixed_reference_domain_residual_time"):
        return mixed_reference_domain_residual_time(c.get("integrand_fun"), c.get("ansatz_fun"), c.get("ref_int_coor"),
                                                    c.get("ref_int_weights"), c.get("mapping_key"))
    return recognise_weak_form(model)


# silence "imported but unused" while keeping the canonical import location for users
_ = spaces
