"""CIFAR-10 provider of the Keras zoo (ref ``keras_model_zoo/data/cifar10.py``) — same
dataset class as the native models use."""
from ...data.cifar10 import Cifar10_data  # noqa: F401
