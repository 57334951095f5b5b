"""Out-of-bounds writes, without compute-sanitizer (closed on this pool): every caller-owned tensor of the step kernels is carved
out of one arena filled with a canary byte, with guard bands between the tensors and ODD byte offsets for the byte tensors (the
kernels emit 16-byte stores at the destination's alignment phase, with ragged heads and tails); after resets and steps at a
ragged batch size every guard byte must be untouched and the outputs must equal those of a batch with ordinary tensors."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
CANARY = 0xA5
GUARD = 512


class Arena(object):
    def __init__(self, nbytes, device):
        self.buf = torch.full((nbytes,), CANARY, dtype=torch.uint8, device=device)
        self.used = torch.zeros((nbytes,), dtype=torch.bool, device=device)
        self.off = GUARD

    def carve(self, like, odd):
        """A tensor shaped like `like` inside the arena; float / int32 tensors keep 16-byte alignment, byte tensors start at an
        odd offset when `odd`."""
        n = like.numel() * like.element_size()
        start = (self.off + 255) // 256 * 256 + (1 + 2 * (self.off % 7) if (odd and like.element_size() == 1) else 0)
        t = self.buf[start:start + n].view(like.dtype).view(like.shape)
        t.copy_(like)
        self.used[start:start + n] = True
        self.off = start + n + GUARD
        assert self.off + GUARD < self.buf.numel()
        return t

    def check(self, what):
        bad = (~self.used) & (self.buf != CANARY)
        assert not bool(bad.any()), "%s: %d guard bytes overwritten, first at arena offset %d" % (what, int(bad.sum()), int(bad.nonzero()[0]))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def test_firemaker_kernel_stays_inside_its_tensors():
    from ai_safety_gridworlds_b200 import _abi
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    N = 16 * 9 + 5                                                     # full batches of 16 games + a ragged one
    ref = FiremakerVectorEnv(N, seed=3, autoreset_mode=1, max_iterations=45)
    env = FiremakerVectorEnv(N, seed=3, autoreset_mode=1, max_iterations=45)
    arena = Arena(64 << 20, env.device)
    names = ("state", "board", "cube", "crop_workers", "crop_supervisor", "lcrop_workers", "lcrop_supervisor", "reward_workers",
             "reward_supervisor", "terminated", "step_type")
    for nm in names:
        setattr(env, nm, arena.carve(getattr(env, nm), odd=nm not in ("state",)))
    env._obs = _abi.GwFmObs(_ptr(env.board), _ptr(env.cube), _ptr(env.crop_workers), _ptr(env.crop_supervisor), _ptr(env.lcrop_workers),
                            _ptr(env.lcrop_supervisor))
    env._out = _abi.GwFmOut(_ptr(env.reward_workers), _ptr(env.reward_supervisor), _ptr(env.terminated), _ptr(env.step_type))
    g = torch.Generator(device=env.device); g.manual_seed(0)
    for t in range(40):
        a = torch.randint(0, 5, (N, 3), dtype=torch.int32, device=env.device, generator=g)
        env.step(a)
        ref.step(a)
        for nm in names[1:]:
            assert torch.equal(getattr(env, nm), getattr(ref, nm)), (nm, t)
    env.reset(torch.arange(N, device=env.device) % 3 == 0)
    torch.cuda.synchronize()
    arena.check("gw_fm_kernel")
    env.close(); ref.close()


@pytest.mark.parametrize("impl", ["tma", "direct"])
def test_single_agent_kernels_stay_inside_their_tensors(impl, monkeypatch):
    from ai_safety_gridworlds_b200 import _abi
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    monkeypatch.setenv("GWSIM_STEP_IMPL", impl)
    for name, kw in (("island_navigation_ex", {}), ("boat_race_ex", {"level": 3})):
        N = 32 * 21 + 7
        ref = VectorEnv(name, N, autoreset_mode=1, **kw)
        env = VectorEnv(name, N, autoreset_mode=1, **kw)
        arena = Arena(32 << 20, env.device)
        names = ("state", "board", "cube", "value_board", "reward", "terminated", "step_type", "reason")
        for nm in names:
            setattr(env, nm, arena.carve(getattr(env, nm), odd=False))   # the TMA bulk stores need the 16-byte aligned slices torch gives
        env._obs = _abi.GwObs(_ptr(env.board), _ptr(env.cube), _ptr(env.value_board))
        env._out = _abi.GwStepOut(_ptr(env.reward), _ptr(env.terminated), _ptr(env.step_type), _ptr(env.reason), None)
        for t in range(60):
            a = ref.random_actions(5, t)
            env.step(a)
            ref.step(a)
            for nm in names[1:]:
                assert torch.equal(getattr(env, nm), getattr(ref, nm)), (name, nm, t)
        env.reset(torch.arange(N, device=env.device) % 2 == 0)
        torch.cuda.synchronize()
        arena.check("%s %s" % (impl, name))
        env.close(); ref.close()
