"""Model assembly with the reference's interface (model.py:34-128).

``Model()`` builds the same sub-modules under the same attribute names (``Transformation``,
``FeatureExtraction``, ``AdaptiveAvgPool``, ``get_semantic_vectors``, ``encoder``, ``decoder``),
so a reference checkpoint (253 tensors in the LSTM/LSTM configuration) loads with
``load_state_dict``.  ``config.ENCODER`` / ``config.DECODER`` select the LSTM path (the north-star hot path) or
the reference's default Transformer / Transformer and Transformer / Linear configurations (modules/transformer.py).  ``Model.forward(input, text, overlap, scene, ious, is_train=True)`` returns
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

# module-level names the reference exposes (model.py:17-32)
character = config.CHARS
converter = AttnLabelConverter(character)
num_classes = len(converter.character)
imgH, imgW = 32, 100          # crop geometry
num_fiducial = 20             # TPS control points
input_channel, output_channel = 1, 512
hidden_size = 256


def _pick(kind: str, key, table):
    """Constructor for the configured variant ``key`` of a model part, or the reference's error."""
    try:
        return table[key]
    except KeyError:
        raise Exception("Model.py %s Error: '%s' not recognized" % (kind, key)) from None


def _semantic_vector_module():
    """config.SEMANTIC_SOURCE wins over config.SEMANTIC_EMBEDDING, like the if-chain of model.py:50-59."""
    by_source = {'zero': Zero, 'rand': Random}
    by_embedding = {'bert': Bert_Embedding, 'linear': Linear_Embedding}
    cls = by_source.get(config.SEMANTIC_SOURCE) or by_embedding.get(config.SEMANTIC_EMBEDDING)
    if cls is None:
        raise Exception("Model.py Semantic vector error: check config.SEMANTIC_SOURCE and config.SEMANTIC_EMBEDDING")
    return cls()


class Model(nn.Module):
    """Sub-modules are created in the reference's order (rectifier, feature extractor, pooling, semantic vectors,
    encoder, decoder): the parameter initialisers then consume the RNG identically, so ``torch.manual_seed(s);
    Model()`` yields the reference's weights (tests/test_cabi_and_host.py)."""

    def __init__(self, imgH=imgH, imgW=imgW, input_channel=input_channel):
        """The reference hard-codes 1x32x100 (model.py:21-32); the keyword arguments exist for the
        3x64x256 stress configuration and default to the reference's values."""
        super().__init__()
        size = (imgH, imgW)
        self.Transformation = TPS_SpatialTransformerNetwork(F=num_fiducial, I_size=size, I_r_size=size,
                                                            I_channel_num=input_channel)
        self.FeatureExtraction = ResNet_FeatureExtractor(input_channel, output_channel)
        self.FeatureExtraction_output = output_channel
        self.AdaptiveAvgPool = nn.AdaptiveAvgPool2d((None, 1))   # attribute parity only: _run averages over H itself
        self.get_semantic_vectors = _semantic_vector_module()
        self.encoder = _pick("Encoder", config.ENCODER, {
            'LSTM': lambda: nn.Sequential(BidirectionalLSTM(output_channel, hidden_size, hidden_size),
                                          BidirectionalLSTM(hidden_size, hidden_size, hidden_size)),
            'Transformer': TF_Encoder,
            'Oscar': Oscar_Bert,
        })()
        self.decoder = _pick("Decoder", config.DECODER, {
            'LSTM': lambda: Attention(hidden_size, hidden_size, num_classes),
            'Transformer': lambda: TF_Decoder(num_classes),
            'Linear': lambda: Linear_Decoder(num_classes),
        })()

    # ---------------------------------------------------------------------------------------
    def _run(self, input, text, overlap, scene, ious, is_train, stages=None):
        lstm_encoder = isinstance(self.encoder, nn.Sequential)
        if input.shape[0] == 0:
            # empty batch (e.g. the last, empty shard of a rank): the reference returns an empty
            # [0, steps, classes] tensor from Attention.forward; nothing is launched here
            if not input.is_cuda:
                raise _cabi.MstrError("the recognition path has no CPU implementation: move the inputs to a CUDA device")
            steps = config.MAX_TEXT_LENGTH + (0 if isinstance(self.decoder, TF_Decoder) and not is_train else 1)
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
        if not lstm_encoder:
            # transformer family (model.py:101-105): the encoder output stays in the activation dtype
            enc = self.encoder.forward_features(visual, semantics) if isinstance(self.encoder, TF_Encoder) \
                else self.encoder.forward_features(visual)
            out = self.decoder.decode(enc, text, is_train, enc_act=enc, semantics=semantics)
            if stages is not None:
                stages.update(rectified=rect, features=feat, visual=visual, semantics=semantics, enc1=ops.cast(enc, torch.float32))
            return out
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
    """``saved_model``: optional path of a checkpoint.  Returns the model wrapped in ``nn.DataParallel`` over
    ``config.DEVICE_IDS`` on ``config.PRIMARY_DEVICE`` like the reference's get_model (model.py:109-128); the wrapper
    is what gives every state-dict key its ``module.`` prefix.  The model is put in eval mode: the kernels implement
    eval-mode BatchNorm only."""
    model = torch.nn.DataParallel(Model(), device_ids=config.DEVICE_IDS).to(config.PRIMARY_DEVICE).eval()
    if not saved_model:
        print('  - Training from scratch (no pretrained weights provided)')
        return model
    print('  - Loading model from:', saved_model)
    state = torch.load(saved_model, map_location='cpu')
    # the reference drops the embedding table of a checkpoint before loading (model.py:121): it is re-initialised
    state.pop('module.get_semantic_vectors.embed.weight', None)
    model.load_state_dict(state, strict=False)
    return model


def load_numpy_state_dict(model: nn.Module, sd) -> nn.Module:
    """Load a {name: numpy array} state dict (e.g. synth.make_state_dict) into ``model``."""
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return model
