"""`from models.lstm import Model` -- the import every reference distill script performs
(LstmDistillFromDinoV2Train.py:5, LstmDistillation.py:5, ...Spampinato.py:5, ...Eval.py:5).  The reference does
not ship this file; this shim resolves it to the B200 encoder."""
from cerebralsignalnetworks_b200.lstm import Model  # noqa: F401
