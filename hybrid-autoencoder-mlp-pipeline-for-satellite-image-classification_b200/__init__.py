"""B200-native supervised autoencoder + MLP hot path (drop-in for the reference notebook's modules).

    from ae_b200 import Encoder, Decoder, SupervisedAutoencoder, MLP, Adam, TrainStep
"""
from . import _lib
from .modules import Encoder, Decoder, SupervisedAutoencoder, MLP, default_backend, default_precision
from .optim import Adam
from .train import TrainStep, MLPTrainStep
from . import dp
from .pipeline import extract_features, encode_predict
from . import data, fit, search
from .data import DeviceDataset, DeviceLoader, TrainTransformAE, EvalTransform, augment_u8

__all__ = ["Encoder", "Decoder", "SupervisedAutoencoder", "MLP", "Adam", "TrainStep", "MLPTrainStep", "dp", "extract_features",
           "encode_predict", "default_backend", "default_precision", "data", "fit", "search", "DeviceDataset", "DeviceLoader",
           "TrainTransformAE", "EvalTransform", "augment_u8"]
