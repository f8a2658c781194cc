"""Host logic of the 0-D observer groups (cherab/tools/observers/group/base.py): membership, broadcast setters, ray bundles."""
import numpy as np
import pytest

import core_b200 as cb


def test_group_membership_and_broadcast_setters():
    group = cb.FibreOpticGroup(name="Divertor Fibre Optic Array")
    for i in range(3):
        group.add_observer(cb.FibreOptic(name=str(i + 1), transform=cb.translate(2.3, 0, 1.25)))
    with pytest.raises(ValueError):
        group.add_observer(cb.SightLine())                       # group/base.py:106-107
    group.acceptance_angle = 1.4
    group.radius = [0.001, 0.002, 0.003]
    group.pixel_samples = 50
    assert group.acceptance_angle == [1.4] * 3 and group.radius == [0.001, 0.002, 0.003] and group.names == ["1", "2", "3"]
    with pytest.raises(ValueError):
        group.radius = [0.001, 0.002]                            # length mismatch
    with pytest.raises(TypeError):
        group.names = "abc"
    group.names = ["a", "b", "c"]
    assert group.observers[1].name == "b"


def test_fibre_ray_bundle_geometry():
    f = cb.FibreOptic(transform=cb.look_at((1.0, 2.0, 3.0), (1.0, 2.0, 0.0), up=(0, 1, 0)), acceptance_angle=10.0, radius=0.01, pixel_samples=500)
    o, d, w = f.rays()
    axis = np.array([0.0, 0.0, -1.0])
    assert o.shape == d.shape == (500, 3)
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-14)
    cos_t = d @ axis
    assert cos_t.min() >= np.cos(np.deg2rad(10.0)) - 1e-12 and np.allclose(cos_t, w)
    assert np.all(np.linalg.norm(o - np.array([1.0, 2.0, 3.0]), axis=1) <= 0.01 + 1e-12)
    np.testing.assert_allclose((o - np.array([1.0, 2.0, 3.0])) @ axis, 0.0, atol=1e-12)        # origins lie in the tip plane
    # uniform in solid angle: the mean of cos(theta) over the cap is (1 + cos_max) / 2
    assert abs(cos_t.mean() - 0.5 * (1.0 + np.cos(np.deg2rad(10.0)))) < 1e-5
    # uniform in area: mean squared radius is R^2 / 2
    r2 = np.sum((o - np.array([1.0, 2.0, 3.0])) ** 2, axis=1)
    assert abs(r2.mean() / 0.01 ** 2 - 0.5) < 2e-3
    assert abs(f.solid_angle - 2 * np.pi * (1 - np.cos(np.deg2rad(10.0)))) < 1e-15


def test_group_transform_applies_to_every_observer():
    g = cb.SightLineGroup([cb.SightLine(transform=cb.translate(1, 0, 0)), cb.SightLine(transform=cb.translate(0, 1, 0))],
                          transform=cb.translate(0, 0, 5))
    o, d, w, owner = g.gather_rays()
    np.testing.assert_allclose(o, [[1, 0, 5], [0, 1, 5]])
    np.testing.assert_allclose(d, [[0, 0, 1], [0, 0, 1]])
    assert list(owner) == [0, 1] and list(w) == [1.0, 1.0]


def test_device_descriptors_of_a_group():
    # what cb2_observer0d_rays_device / _reduce_device receive: one cb2_observer0d per observer (group transform applied), the rays'
    # offsets and the etendue (solid angle x collection area for a fibre, the sensitivity for a sight line)
    g = cb.FibreOpticGroup(transform=cb.translate(0.0, 0.0, 1.0))
    g.add_observer(cb.FibreOptic(transform=cb.translate(1.0, 2.0, 3.0), acceptance_angle=10.0, radius=0.002, pixel_samples=7))
    g.add_observer(cb.FibreOptic(transform=cb.translate(-1.0, 0.0, 0.0), acceptance_angle=5.0, radius=0.001, pixel_samples=3))
    arr, offs, etendue = g._descs()
    assert list(offs) == [0, 7, 10]
    assert [arr[i].samples for i in range(2)] == [7, 3] and arr[0].radius == 0.002 and arr[1].acceptance_angle == 5.0
    assert (arr[0].to_world[3], arr[0].to_world[7], arr[0].to_world[11]) == (1.0, 2.0, 4.0)            # group translation on top
    ob = g.observers[0]
    np.testing.assert_allclose(etendue[0], 2 * np.pi * (1 - np.cos(np.deg2rad(10.0))) * np.pi * 0.002 ** 2, rtol=1e-14)
    assert etendue[0] == ob.solid_angle * ob.collection_area
    s = cb.SightLineGroup([cb.SightLine(sensitivity=2.5), cb.SightLine()])
    arr, offs, etendue = s._descs()
    assert list(offs) == [0, 1, 2] and list(etendue) == [2.5, 1.0] and arr[0].radius == 0.0 and arr[1].samples == 1
    g.acceptance_angle = 95.0
    with pytest.raises(ValueError):
        g._descs()


def test_group_pipelines_api():
    # group/base.py:383-436: pipelines as a list of lists, connect_pipelines builds a fresh set per observer
    g = cb.FibreOpticGroup([cb.FibreOptic(), cb.FibreOptic(), cb.FibreOptic()])
    assert g.pipelines == [[], [], []]
    g.connect_pipelines([cb.SpectralRadiancePipeline0D, cb.SpectralPowerPipeline0D], [{"name": "MySpectralPipeline"}, {}])
    assert [[type(p).__name__ for p in ps] for ps in g.pipelines] == [["SpectralRadiancePipeline0D", "SpectralPowerPipeline0D"]] * 3
    assert g.pipelines[0][0] is not g.pipelines[1][0] and g.pipelines[2][0].name == "MySpectralPipeline"
    assert all(p.display_progress is False for ps in g.pipelines for p in ps)
    with pytest.raises(ValueError):
        g.connect_pipelines([cb.RadiancePipeline0D], [{}, {}])
    with pytest.raises(ValueError):
        g.pipelines = [[cb.RadiancePipeline0D()]]
    g.pipelines = [[cb.RadiancePipeline0D()], [], [cb.PowerPipeline0D()]]
    assert len(g.observers[0].pipelines) == 1 and g.observers[1].pipelines == []
