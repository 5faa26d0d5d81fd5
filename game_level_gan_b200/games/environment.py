"""Interface shared by the batched environments (mirrors games/environment.py:4-44 of the reference)."""
import abc


class MultiEnvironment(abc.ABC):
    """A batch of independent multi-player games stepped in lock-step."""

    @abc.abstractmethod
    def players_layer_shape(self):
        """Shape of the per-player layer of the observation (grid games)."""

    @abc.abstractmethod
    def state_shape(self):
        """Shape of one player's observation of one game."""

    @abc.abstractmethod
    def reset(self, *args, **kwargs):
        """Start a new episode on the given levels; returns the initial observations."""

    @abc.abstractmethod
    def step(self, actions):
        """Apply one action per player and game; returns (observations, rewards)."""

    @property
    @abc.abstractmethod
    def actions(self):
        """Number of discrete actions."""
