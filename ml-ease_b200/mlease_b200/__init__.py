"""mlease_b200 -- B200-native ADMM logistic regression behind ml-ease's AdmmTrain/NaiveTrain/Test surface."""
from ._native import MleaseError, SO_PATH, lib  # noqa: F401
from .admm import AdmmSession, Comm, World, naive_train, naive_train_dense, score, test_loglik  # noqa: F401
