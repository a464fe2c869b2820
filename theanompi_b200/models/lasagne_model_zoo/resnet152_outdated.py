"""ResNet-152 (ref ``lasagne_model_zoo/resnet152_outdated.py``): same network family with
block counts [3, 8, 36, 3]."""
from .resnet50 import ResNet50


class ResNet152(ResNet50):
    blocks = (3, 8, 36, 3)

    def __init__(self, config):
        super().__init__(config)
        self.name = "ResNet152"
