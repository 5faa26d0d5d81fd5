"""CaptureSafePolicy (games/rollout.py) keeps a reference-style recurrent policy's state in fixed buffers; the numbers
must be those of the bare network (policies/LSTMPolicy.py:26-41 rebinds `self.state` on every forward)."""
import torch


class RebindingPolicy(torch.nn.Module):
    """same state handling as the reference's LSTMPolicy: `state` is None after reset and rebound on every call"""

    def __init__(self):
        super().__init__()
        self.cells = torch.nn.ModuleList([torch.nn.LSTMCell(5, 7), torch.nn.LSTMCell(7, 7)])
        self.head = torch.nn.Linear(7, 3)
        self.state = None

    def reset_state(self):
        self.state = None

    def forward(self, x):
        if self.state is None:
            self.state = [(x.new_zeros(x.size(0), 7), x.new_zeros(x.size(0), 7)) for _ in self.cells]
        h = x
        for i, cell in enumerate(self.cells):
            self.state[i] = cell(h, self.state[i])
            h = self.state[i][0]
        return self.head(h), h.sum(-1)


def test_wrapped_policy_equals_the_bare_one_and_keeps_its_buffers():
    from game_level_gan_b200.games.rollout import CaptureSafePolicy
    torch.manual_seed(0)
    bare = RebindingPolicy()
    twin = RebindingPolicy()
    twin.load_state_dict(bare.state_dict())
    safe = CaptureSafePolicy(twin)
    xs = torch.randn(9, 4, 5)
    with torch.no_grad():
        for episode in range(2):
            bare.reset_state()
            safe.reset_state()
            ptrs = None
            for x in xs:
                a, b = bare(x), safe(x)
                assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
                now = [t.data_ptr() for hc in safe._buf for t in hc]
                assert ptrs is None or ptrs == now            # the recurrent state never moves
                ptrs = now
    assert set(safe.state_dict().keys()) == set(bare.state_dict().keys())
