"""`games` package of the reference, B200-native (games/__init__.py:1-4)."""
from .environment import MultiEnvironment
from .race import Race, RaceCar
from .race_utils import RaceConfig, predefined_tracks, race_game

__all__ = ['MultiEnvironment', 'Race', 'RaceCar', 'RaceConfig', 'predefined_tracks', 'race_game']
