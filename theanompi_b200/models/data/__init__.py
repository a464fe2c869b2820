from .cifar10 import Cifar10_data  # noqa: F401
from .imagenet import ImageNet_data  # noqa: F401
from .mnist import MNIST_data  # noqa: F401
