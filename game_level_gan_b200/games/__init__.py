"""`games` package of the reference, B200-native (games/__init__.py:1-4)."""
from .environment import MultiEnvironment
from .pacman import Pacman
from .race import Race, RaceCar
from .race_utils import RaceConfig, predefined_tracks, race_game
from .pytorch_wrapper import PytorchWrapper
from . import game_helpers
from .rollout import CaptureSafePolicy, GraphedRollout, HostRollout, HostStepper, capture_safe_agents

__all__ = ['MultiEnvironment', 'Pacman', 'Race', 'RaceCar', 'RaceConfig', 'predefined_tracks', 'race_game',
           'PytorchWrapper', 'game_helpers', 'GraphedRollout', 'HostStepper', 'HostRollout', 'CaptureSafePolicy',
           'capture_safe_agents']
