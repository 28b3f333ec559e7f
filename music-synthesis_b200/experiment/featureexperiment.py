"""The stage-one caller of this path (featuresynth/experiment/featureexperiment.py:77-100): a
vocoder turns generated feature sequences into audio.  `NeuralVocoder` wraps any of the
generators here (the stage-one experiments use it at 512 frames); the feature generators
themselves are out of scope."""
import torch


class BaseVocoder(object):
    def __call__(self, features):
        raise NotImplementedError()


class NeuralVocoder(BaseVocoder):
    def __init__(self, network):
        self.network = network

    def __call__(self, features):
        with torch.no_grad():
            return self.network(features)
