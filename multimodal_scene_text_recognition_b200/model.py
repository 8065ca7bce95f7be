"""Model assembly with the reference's interface (model.py:34-128).

``Model()`` builds the same sub-modules under the same attribute names (``Transformation``,
``FeatureExtraction``, ``AdaptiveAvgPool``, ``get_semantic_vectors``, ``encoder``, ``decoder``),
so a reference checkpoint (253 tensors in the LSTM/LSTM configuration) loads with
``load_state_dict``.  ``Model.forward(input, text, overlap, scene, ious, is_train=True)`` returns
the ``[B, MAX_TEXT_LENGTH + 1, num_classes]`` logits like the reference; ``recognize`` is the same
pass returning the decoded ids and lengths the fused decoder already produced.

Inference only: the kernels have no backward.  There is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi, config, ops
from ._pack import act_dtype
from .utils import AttnLabelConverter
from .modules.transformation import TPS_SpatialTransformerNetwork
from .modules.feature_extraction import ResNet_FeatureExtractor
from .modules.semantic_vectors import Linear_Embedding, Bert_Embedding, Zero, Random
from .modules.encoders import BidirectionalLSTM, TF_Encoder, Oscar_Bert
from .modules.decoders import Attention, TF_Decoder, Linear_Decoder

character = config.CHARS
converter = AttnLabelConverter(character)
num_classes = len(converter.character)

# model.py:21-32
imgH = 32
imgW = 100
num_fiducial = 20
input_channel = 1
output_channel = 512
hidden_size = 256


class Model(nn.Module):

    def __init__(self, imgH=imgH, imgW=imgW, input_channel=input_channel):
        """The reference hard-codes 1x32x100 (model.py:21-32); the keyword arguments exist for the
        3x64x256 stress configuration and default to the reference's values."""
        super(Model, self).__init__()
        self.Transformation = TPS_SpatialTransformerNetwork(
            F=num_fiducial, I_size=(imgH, imgW), I_r_size=(imgH, imgW), I_channel_num=input_channel)

        self.FeatureExtraction = ResNet_FeatureExtractor(input_channel, output_channel)
        self.FeatureExtraction_output = output_channel
        self.AdaptiveAvgPool = nn.AdaptiveAvgPool2d((None, 1))   # kept for attribute parity; see forward

        if config.SEMANTIC_SOURCE == 'zero':
            self.get_semantic_vectors = Zero()
        elif config.SEMANTIC_SOURCE == 'rand':
            self.get_semantic_vectors = Random()
        elif config.SEMANTIC_EMBEDDING == 'bert':
            self.get_semantic_vectors = Bert_Embedding()
        elif config.SEMANTIC_EMBEDDING == 'linear':
            self.get_semantic_vectors = Linear_Embedding()
        else:
            raise Exception("Model.py Semantic vector error: check config.SEMANTIC_SOURCE and config.SEMANTIC_EMBEDDING")

        if config.ENCODER == 'LSTM':
            self.encoder = nn.Sequential(
                BidirectionalLSTM(self.FeatureExtraction_output, hidden_size, hidden_size),
                BidirectionalLSTM(hidden_size, hidden_size, hidden_size))
        elif config.ENCODER == 'Transformer':
            self.encoder = TF_Encoder()
        elif config.ENCODER == 'Oscar':
            self.encoder = Oscar_Bert()
        else:
            raise Exception("Model.py Encoder Error: '" + config.ENCODER + "' not recognized")

        if config.DECODER == "LSTM":
            self.decoder = Attention(256, 256, num_classes)
        elif config.DECODER == "Transformer":
            self.decoder = TF_Decoder(num_classes)
        elif config.DECODER == "Linear":
            self.decoder = Linear_Decoder(num_classes)
        else:
            raise Exception("Model.py Decoder Error: '" + config.DECODER + "' not recognized")

    # ---------------------------------------------------------------------------------------
    def _run(self, input, text, overlap, scene, ious, is_train, stages=None):
        if config.ENCODER != 'LSTM':
            raise NotImplementedError("only config.ENCODER == 'LSTM' is implemented (model.py:99)")
        if input.shape[0] == 0:
            # empty batch (e.g. the last, empty shard of a rank): the reference returns an empty
            # [0, steps, classes] tensor from Attention.forward; nothing is launched here
            if not input.is_cuda:
                raise _cabi.MstrError("the recognition path has no CPU implementation: move the inputs to a CUDA device")
            steps = config.MAX_TEXT_LENGTH + 1
            dev = input.device
            return {"logits": torch.zeros((0, steps, self.decoder.num_classes), device=dev),
                    "ids": torch.zeros((0, steps), device=dev, dtype=torch.int32),
                    "lengths": torch.zeros((0,), device=dev, dtype=torch.int32)}
        # Transformation (model.py:88)
        rect = self.Transformation(input)
        # Feature extraction (model.py:91-93): NHWC result; permute + AdaptiveAvgPool2d((None,1)) +
        # squeeze is the mean over H of the NHWC tensor
        feat = self.FeatureExtraction.forward_nhwc(rect)
        # [B, W', C] activation dtype; with H' == 1 (the reference geometry) the mean is the tensor itself
        visual = feat.view(feat.shape[0], feat.shape[2], feat.shape[3]) if feat.shape[1] == 1 else ops.mean_h(feat)
        # Semantic vectors (model.py:96); computed because the reference computes them - the LSTM
        # encoder / attention decoder never read them (decoders.py:79-101)
        semantics = self.get_semantic_vectors(overlap, scene, ious)
        # Encode (model.py:100)
        bf16 = config.PRECISION == 'bf16'
        enc0 = self.encoder[0].forward_features(visual, want_f32=not bf16, want_act=bf16)
        # the tensor-core decoder consumes the bf16 encoder output only; the f32 copy is produced for the
        # fp32 SIMT decoder and for stage dumps
        need_f32 = (not bf16) or (not config.RNN_TENSOR_CORES) or stages is not None
        enc1 = self.encoder[1].forward_features(enc0, want_f32=need_f32, want_act=bf16)
        if bf16:
            enc1_act, enc1_f32 = enc1 if need_f32 else (enc1, None)     # ops.linear returns (bf16, f32)
        else:
            enc1_act, enc1_f32 = None, enc1
        # Decode (model.py:105)
        out = self.decoder.decode(enc1_f32 if enc1_f32 is not None else enc1_act, text, is_train, enc_act=enc1_act)
        if stages is not None:
            stages.update(rectified=rect, features=feat, visual=visual, semantics=semantics,
                          enc0=enc0, enc1=enc1_f32)
        return out

    def forward(self, input, text, overlap, scene, ious, is_train=True):
        return self._run(input, text, overlap, scene, ious, is_train)["logits"]

    def recognize(self, input, overlap, scene, ious, stages=None):
        """Greedy recognition: dict(logits [B,S,C] f32, ids [B,S] i32, lengths [B] i32)."""
        return self._run(input, None, overlap, scene, ious, False, stages)


def get_model(saved_model=None):
    '''
    input: saved_model = the path to pretrained weights to load model from, if None then wont load any weights
    return: model  (nn.DataParallel-wrapped on config.PRIMARY_DEVICE, like model.py:109-128)
    '''
    model = Model()
    model = torch.nn.DataParallel(model, device_ids=config.DEVICE_IDS)
    model = model.to(config.PRIMARY_DEVICE)
    model.eval()   # the kernels implement eval-mode BatchNorm only

    if saved_model:
        print('  - Loading model from:', saved_model)
        pretrained_dict = torch.load(saved_model, map_location=lambda storage, loc: storage)
        pretrained_dict.pop('module.get_semantic_vectors.embed.weight', None)   # model.py:121
        model.load_state_dict(pretrained_dict, strict=False)
    else:
        print('  - Training from scratch (no pretrained weights provided)')

    return model


def load_numpy_state_dict(model: nn.Module, sd) -> nn.Module:
    """Load a {name: numpy array} state dict (e.g. synth.make_state_dict) into ``model``."""
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return model
