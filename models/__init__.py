"""Drop-in for the reference's (missing) `models` package: `from models.lstm import Model`."""
