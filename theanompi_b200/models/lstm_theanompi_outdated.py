"""Import-path parity with the reference's ``lstm_theanompi_outdated.py`` (the model-contract
adapter); the maintained implementation lives in :mod:`theanompi_b200.models.lstm`."""
from .lstm import IMDB_Data, LSTM  # noqa: F401
