"""The RGB observation: the colour tables read out of the reference (envs/colours.json) reproduce the
recorded `RGB` arrays of the reference's distiller on the CPU, and the CUDA look-up kernel (gw_render_rgb, through the C ABI)
and GridworldGymEnv.render('rgb_array') reproduce them on the B200."""
import numpy as np
import pytest

from conftest import load_golden, rgb_golden_names


@pytest.mark.parametrize("name", rgb_golden_names())
def test_colour_tables_reproduce_the_reference_rgb(name):
    from ai_safety_gridworlds_b200 import render
    d, meta = load_golden(name)
    lut = render.rgb_lut(meta["env"])
    got = np.moveaxis(lut[d["board"]], -1, 1)              # [T, H, W, 3] -> [T, 3, H, W]
    np.testing.assert_array_equal(got, d["rgb"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", rgb_golden_names())
def test_cuda_rgb_kernel_matches_the_reference(name):
    torch = pytest.importorskip("torch")
    from ai_safety_gridworlds_b200 import render
    d, meta = load_golden(name)
    dev = torch.device("cuda", 0)
    lut = torch.from_numpy(render.rgb_lut(meta["env"])).to(dev)
    boards = torch.from_numpy(d["board"]).to(dev)
    np.testing.assert_array_equal(render.render_rgb(boards, lut).cpu().numpy(), d["rgb"])
    # padded rows (the savanna / sokoban tensors are views of 16-byte pitched buffers)
    T, H, W = d["board"].shape
    buf = torch.zeros((T, H * W + 13), dtype=torch.uint8, device=dev)
    buf[:, :H * W] = boards.reshape(T, -1)
    view = buf[:, :H * W].unflatten(-1, (H, W))
    np.testing.assert_array_equal(render.render_rgb(view, lut).cpu().numpy(), d["rgb"])


@pytest.mark.gpu
def test_gym_render_rgb_array():
    torch = pytest.importorskip("torch")
    from ai_safety_gridworlds_b200 import render
    from ai_safety_gridworlds_b200.helpers.gridworld_gym_env import GridworldGymEnv
    for name, kw in (("island_navigation_ex", {}), ("whisky_gold", {}), ("side_effects_sokoban", {"level": 1})):
        env = GridworldGymEnv(name, **kw)
        env.reset()
        for t in range(5):
            env.step(1 + t % 4)
        rgb = env.render("rgb_array")
        board = np.array([[ord(ch) for ch in row[::2]] for row in env.render("ansi").split("\n")], np.uint8)   # characters joined by blanks
        assert rgb.dtype == np.uint8 and rgb.shape == (3,) + board.shape
        np.testing.assert_array_equal(rgb, np.moveaxis(render.rgb_lut(name)[board], -1, 0))
    env = GridworldGymEnv("boat_race_ex", level=3, num_envs=257)
    env.reset()
    rgb = env.render("rgb_array")
    assert tuple(rgb.shape) == (257, 3, 7, 7) and rgb.is_cuda
    lut = render.rgb_lut("boat_race_ex")
    np.testing.assert_array_equal(rgb.cpu().numpy(), np.moveaxis(lut[env.vector_env.board.cpu().numpy()], -1, 1))
