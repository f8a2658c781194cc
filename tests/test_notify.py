"""Scene plumbing: Notifier semantics (cherab/core/utility/notify.py:33-162) and the change
notifications of Plasma / Composition / ModelManager (cherab/core/plasma/node.pyx:33-198, 323-331, 545-554)."""
import gc

import numpy as np
import pytest

import core_b200 as cb
from core_b200.slab import build_constant_slab_plasma


class Counter:
    def __init__(self):
        self.n = 0

    def hit(self):
        self.n += 1


def test_notifier_functions_methods_and_weak_references():
    n = cb.Notifier()
    c = Counter()
    calls = []

    def fn():
        calls.append(1)

    n.add(c.hit)
    n.add(c.hit)                      # registering twice has no effect
    n.add(fn)
    assert n.is_present(c.hit) and n.is_present(fn)
    n.notify()
    assert c.n == 1 and len(calls) == 1
    n.remove(fn)
    n.notify()
    assert c.n == 2 and len(calls) == 1
    del c                              # weak reference: the observer may die, the notifier purges it
    gc.collect()
    n.notify()
    assert n._callbacks_refs == []


def test_plasma_changes_notify():
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=100.,
                                        plasma_species=[(cb.deuterium, 1, 1e19, 100., (0, 0, 0))])
    c = Counter()
    plasma.notifier.add(c.hit)
    plasma.models = [cb.Bremsstrahlung()]
    assert c.n == 1
    plasma.models.add(cb.Bremsstrahlung())
    assert c.n == 2 and len(plasma.models) == 2
    sp = plasma.composition.get(cb.deuterium, 1)
    plasma.composition.add(cb.Species(cb.deuterium, 0, sp.distribution))
    assert c.n == 3 and len(plasma.composition) == 2
    plasma.integrator = cb.NumericalIntegrator(step=0.01)
    plasma.b_field = (0, 1.0, 0)
    plasma.atomic_data = cb.AtomicData()
    assert c.n == 6
    plasma.models.clear()
    plasma.composition.clear()
    assert c.n == 8 and len(plasma.models) == 0 and len(plasma.composition) == 0
    with pytest.raises(TypeError):
        plasma.models = [object()]
    with pytest.raises(TypeError):
        plasma.composition = [object()]


@pytest.mark.gpu
def test_renderer_rebuilds_only_after_a_change():
    from core_b200.engine import PlasmaRenderer
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=2000.,
                                        plasma_species=[(cb.deuterium, 1, 1e19, 2000., (0, 0, 0))])
    plasma.atomic_data = cb.AtomicData()
    plasma.models = [cb.Bremsstrahlung()]
    rays = cb.ray_segments(plasma.geometry, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    r = PlasmaRenderer(plasma, 400., 800., 64)
    a, _ = r.render(rays)
    b, _ = r.render(rays)
    assert r.rebuilds == 1 and np.array_equal(a, b)
    # doubling the ion density doubles the continuum (bremsstrahlung.pyx:79-88) — but only if the device scene is rebuilt
    sp = plasma.composition.get(cb.deuterium, 1)
    d = sp.distribution
    plasma.composition.add(cb.Species(cb.deuterium, 1, cb.Maxwellian(cb.Constant3D(2e19), d.temperature, d.velocity, d.atomic_mass)))
    c, _ = r.render(rays)
    assert r.rebuilds == 2
    assert np.allclose(c, 2.0 * a, rtol=1e-6)
    r.close()
