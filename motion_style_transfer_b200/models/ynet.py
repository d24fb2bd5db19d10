"""Drop-in ``YNet`` for the reference's models/ynet.py (Y-Net + MoSA / Y-Net-Mod).

Same constructor, attributes, method set and -- module for module -- the same parameter names,
shapes, creation order (hence identical default initialisation under a given seed) as
/root/reference/models/ynet.py:474-600, so ``state_dict`` files are interchangeable.  The modules
are parameter containers: all arithmetic is done by ``YNetEngine`` through libynet_b200.so.

Supported on the hot path: ``network in {'original', 'fusion'}`` with plain convs or MoSA/LoRA
convs (``train_net`` containing ``mosa``; ``position`` = stage ids or scene/motion/fusion).

SURVEY 8f rank 3 (the paper's comparison baselines, ynet.py:15-131,237-283): the serial / parallel
adapters, layer level (``AdapterLayer``, ``train_net`` containing ``Layer``, e.g. ``parallelLayer_3x3`` of
scripts/sdd/ped_to_biker/tune_pa.sh:22) and block level (``AdapterBlock`` inside ``YNetEncoderB``), are
parameter containers with the reference's names; the engines fold them into the conv they decorate
(inference: exact; BatchNorm in eval mode is an affine map) or run them as one extra conv launch.
Fine-tuning is supported for the parallel layer adapters (linear in weight space); serial adapters
train with batch statistics and raise in training mode.
The ``embed`` network (three conv + ReLU layers on the semantic map and on the observed maps before the encoder,
ynet.py:154-167,529-531) runs through the float32 conv kernel, forward and backward.  The ``semantic`` adapter
(ynet.py:513-519) cannot be constructed in the reference itself (TypeError) and raises the same error here.
"""
import os

import torch
import torch.nn as nn

from . import lora
from .. import ops
from ..engine import YNetEngine, YNetEngineSplit, YNetEngineTC, ChannelCat
from ..utils.softargmax import SoftArgmax2D


def get_conv2d(train_net, l, position, kernel_size, in_channels, out_channels=None, rank=None, stride=1,
               padding=None):
    """Adapter factory (ynet.py:134-151): LoRA conv on adapted positions, plain conv elsewhere."""
    out_channels = in_channels if out_channels is None else out_channels
    padding = kernel_size // 2 if padding is None else padding
    pos = [str(i) for i in (position or [])]
    if 'mosa' in train_net and str(l) in pos:
        assert rank != 0 and rank is not None
        return lora.Conv2d(in_channels, out_channels, kernel_size=kernel_size, r=rank, stride=stride,
                           padding=padding)
    if 'Layer' in train_net and str(l) in pos:
        return AdapterLayer(in_channels, out_channels, kernel_size, adapter_name=train_net, stride=stride,
                            padding=padding)
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)


def _plain_conv(in_channels, out_channels=None, kernel_size=1, stride=1, is_bias=False):
    """ynet.py:8-12: the adapters' own convs (same padding, no bias by default)."""
    out_channels = in_channels if out_channels is None else out_channels
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2,
                     bias=is_bias)


def _adapter_sizes(adapter_name):
    """'parallelLayer_1x1_3x3' -> ['1x1', '3x3'] (ynet.py:21-22)."""
    return adapter_name.split('_')[1:]


def _build_adapter(owner, adapter_name, in_channels, out_channels, serial_channels, stride, is_bias):
    """Submodules of an adapter with the reference's names and creation order (ynet.py:24-50, 88-115): the default
    initialisation consumes the global RNG exactly like the reference before the weights are zeroed."""
    sizes = _adapter_sizes(adapter_name)
    if 'serial' in adapter_name:
        owner.serial_layer = nn.Sequential(nn.BatchNorm2d(serial_channels), _plain_conv(serial_channels, is_bias=is_bias))
        nn.init.zeros_(owner.serial_layer[1].weight)
        if is_bias:
            nn.init.zeros_(owner.serial_layer[1].bias)
    elif 'parallel' in adapter_name and len(sizes) < 2:
        k = int(sizes[0].split('x')[0]) if sizes else 1
        owner.parallel_layer = _plain_conv(in_channels, out_channels, k, stride, is_bias)
        for p in owner.parallel_layer.parameters():
            nn.init.zeros_(p)
    elif 'parallel' in adapter_name:
        owner.parallel_layer = nn.ModuleList(
            [_plain_conv(in_channels, out_channels, int(s.split('x')[0]), stride, is_bias) for s in sizes])
        for p in owner.parallel_layer.parameters():
            nn.init.zeros_(p)
    else:
        raise ValueError(f'Invalid adapter={adapter_name}')


class AdapterBlock(nn.Module):
    """Block-level adapter of YNetEncoderB (ynet.py:15-66): serial = x + conv1x1(BN(x)) on a stage output,
    parallel = sum of k x k convs of the stage INPUT added to the stage output.  Parameter container."""

    def __init__(self, adapter_name, in_channels, out_channels=None, stride=1, is_bias=False):
        super().__init__()
        self.is_bias = is_bias
        self.adapter_name = adapter_name
        self.adapter_size = _adapter_sizes(adapter_name)
        self.is_multiple = len(self.adapter_size) >= 2
        _build_adapter(self, adapter_name, in_channels, out_channels, in_channels, stride, is_bias)


class AdapterLayer(nn.Conv2d):
    """Layer-level adapter (ynet.py:69-131): a conv whose output gets + conv1x1(BN(out)) (serial) or + sum of k x k
    convs of the same input (parallel).  Parameter container; the engines fold it into ONE 3x3 conv."""

    def __init__(self, in_channels, out_channels, kernel_size, adapter_name, adapter_dropout=0., stride=1,
                 is_bias=False, **kwargs):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, **kwargs)
        self.is_bias = is_bias
        self.adapter_name = adapter_name
        self.adapter_size = _adapter_sizes(adapter_name)
        self.is_multiple = len(self.adapter_size) >= 2
        _build_adapter(self, adapter_name, in_channels, out_channels, out_channels, stride, is_bias)


def _mosa_rank(train_net):
    if 'mosa' not in train_net:
        return None
    parts = train_net.split('_')
    return int(parts[1]) if len(parts) > 1 else 1


def _pool():
    return nn.MaxPool2d(kernel_size=2, stride=2, padding=0, dilation=1, ceil_mode=False)


def _double_conv_stage(train_net, l, position, cin, cout, rank):
    return nn.Sequential(
        _pool(),
        get_conv2d(train_net, l, position, 3, cin, cout, rank), nn.ReLU(inplace=False),
        get_conv2d(train_net, l, position, 3, cout, cout, rank), nn.ReLU(inplace=False))


def _conv3x3_f32(x, weight, bias, relu):
    """One 3x3 conv (+ReLU) of a float32 NCHW map through the CUDA-core kernel; differentiable when grads are on."""
    x = x.float().contiguous()
    N, _, H, W = x.shape
    if torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad or (bias is not None and bias.requires_grad)):
        from .. import autograd_engine
        return autograd_engine.Conv3x3Fn.apply(weight, bias, None, None, relu, (ops.SRC_DIRECT,), H, W, x)
    packed = ops.lora_fold(weight.detach().contiguous(), None, None, packed=True)
    return ops.conv3x3_f32([(x, ops.SRC_DIRECT)], packed, None if bias is None else bias.detach(), relu, N, H, W)


class Embedding(nn.Module):
    """ynet.py:154-167: three (conv3x3 + ReLU) layers, channels -> channels (``network='embed'``)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=False),
            nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=False),
            nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=False))

    def forward(self, x):
        for m in self.conv:
            if isinstance(m, nn.Conv2d):
                x = _conv3x3_f32(x, m.weight, m.bias, True)
        return x


class YNetEncoder(nn.Module):
    """ynet.py:170-215: stage 0 = conv+ReLU; stages 1..n-1 = pool, 2 x (conv+ReLU); last = pool."""

    def __init__(self, in_channels, channels=(64, 128, 256, 512, 512), train_net=None, position=[]):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = channels
        self.train_net = train_net
        self.position = position
        self.rank = _mosa_rank(train_net)
        self.stages = nn.ModuleList()
        self.stages.append(nn.Sequential(
            get_conv2d(train_net, 0, position, 3, in_channels, channels[0], self.rank), nn.ReLU(inplace=False)))
        for i in range(len(channels) - 1):
            self.stages.append(_double_conv_stage(train_net, i + 1, position, channels[i], channels[i + 1], self.rank))
        self.stages.append(nn.Sequential(_pool()))


class YNetEncoderL(YNetEncoder):
    pass


class YNetEncoderB(YNetEncoder):
    """ynet.py:237-283: the plain encoder plus, for ``train_net`` containing serial / parallel, one AdapterBlock per
    adapted stage (``adapters[j]`` belongs to stage ``position[j]``)."""

    def __init__(self, in_channels, channels=(64, 128, 256, 512, 512), train_net=None, position=[]):
        pos = []
        for i in position:
            try:
                pos.append(int(i))
            except (TypeError, ValueError):
                pos.append(i)
        super().__init__(in_channels, channels, train_net, pos)
        par_channels_in = [in_channels] + list(channels[:-1])
        if 'serial' in train_net:
            self.adapters = nn.ModuleList([AdapterBlock(train_net, channels[i]) for i in self.position])
        elif 'parallel' in train_net:
            self.adapters = nn.ModuleList(
                [AdapterBlock(train_net, par_channels_in[i], channels[i]) for i in self.position])


class YNetEncoderFusion(nn.Module):
    """Y-Net-Mod encoder (ynet.py:286-395): separate scene / motion branches, then fusion stages."""

    def __init__(self, scene_channel, motion_channel, channels, train_net=None, position=[], n_fusion=2):
        super().__init__()
        self.scene_channel = scene_channel
        self.motion_channel = motion_channel
        self.channels = channels
        self.train_net = train_net
        self.position = position
        self.rank = _mosa_rank(train_net)
        assert not any([i % 2 for i in channels]), f'Odd value in channels={channels}'
        assert n_fusion <= len(channels) - 1, 'The number of fusion exceeds the total number of layer in encoder'
        r = self.rank
        self.scene_stages = nn.ModuleList([nn.Sequential(
            get_conv2d(train_net, 'scene', position, 3, scene_channel, channels[0] // 2, r), nn.ReLU(inplace=False))])
        self.motion_stages = nn.ModuleList([nn.Sequential(
            get_conv2d(train_net, 'motion', position, 3, motion_channel, channels[0] // 2, r), nn.ReLU(inplace=False))])
        self.fusion_stages = nn.ModuleList()
        n_sep = len(channels) - n_fusion - 1
        for i in range(n_sep):
            self.scene_stages.append(
                _double_conv_stage(train_net, 'scene', position, channels[i] // 2, channels[i + 1] // 2, r))
        for i in range(n_sep):
            self.motion_stages.append(
                _double_conv_stage(train_net, 'motion', position, channels[i] // 2, channels[i + 1] // 2, r))
        for i in range(n_sep, len(channels) - 1):
            self.fusion_stages.append(
                _double_conv_stage(train_net, 'fusion', position, channels[i], channels[i + 1], r))
        self.fusion_stages.append(nn.Sequential(_pool()))


class YNetDecoder(nn.Module):
    """ynet.py:398-451: center (2 convs), 5 x [bilinear x2, upsample_conv, cat skip, 2 convs], 1x1 predictor."""

    def __init__(self, encoder_channels, decoder_channels, output_len, traj=False):
        super().__init__()
        if traj:
            encoder_channels = [c + traj for c in encoder_channels]
        encoder_channels = encoder_channels[::-1]
        center = encoder_channels[0]

        def c3(i, o):
            return nn.Conv2d(i, o, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1))

        self.center = nn.Sequential(c3(center, center * 2), nn.ReLU(inplace=False),
                                    c3(center * 2, center * 2), nn.ReLU(inplace=False))
        up_in = [center * 2] + decoder_channels[:-1]
        up_out = [c // 2 for c in up_in]
        self.upsample_conv = nn.ModuleList([c3(i, o) for i, o in zip(up_in, up_out)])
        dec_in = [e + d for e, d in zip(encoder_channels, up_out)]
        self.decoder = nn.ModuleList([
            nn.Sequential(c3(i, o), nn.ReLU(inplace=False), c3(o, o), nn.ReLU(inplace=False))
            for i, o in zip(dec_in, decoder_channels)])
        self.predictor = nn.Conv2d(in_channels=decoder_channels[-1], out_channels=output_len, kernel_size=1,
                                   stride=1, padding=0)


class YNet(nn.Module):
    def __init__(self, obs_len, pred_len, segmentation_model_fp, use_features_only=False, n_semantic_classes=6,
                 encoder_channels=[], decoder_channels=[], n_waypoints=1, train_net=None, position=[],
                 network=None, n_fusion=None):
        super().__init__()
        self.train_net = train_net
        if segmentation_model_fp is not None:
            # third-party pickled smp U-Net, runs once per scene (ynet.py:495-507): kept as a torch module
            map_location = None if torch.cuda.is_available() else torch.device('cpu')
            self.semantic_segmentation = torch.load(segmentation_model_fp, map_location=map_location,
                                                    weights_only=False)
            if use_features_only:
                self.semantic_segmentation.segmentation_head = nn.Identity()
                n_semantic_classes = 16
        else:
            self.semantic_segmentation = nn.Identity()
        self.feature_channels = n_semantic_classes + obs_len
        self.network = network
        if 'semantic' in train_net:
            # ynet.py:513-519 passes position=None into get_conv2d, which iterates over it (ynet.py:140): the reference
            # cannot construct this model (verified against the live reference), so there is no behaviour to reproduce
            raise TypeError("train_net containing 'semantic': the reference's semantic adapter cannot be constructed "
                            "(ynet.py:516 -> ynet.py:140: 'NoneType' object is not iterable)")
        if network == 'fusion':
            assert n_fusion is not None
            self.encoder = YNetEncoderFusion(n_semantic_classes, obs_len, encoder_channels, train_net=train_net,
                                             position=position, n_fusion=n_fusion)
        elif network == 'original' or network == 'embed':
            if network == 'embed':            # ynet.py:529-531
                self.scene_embedding = Embedding(n_semantic_classes)
                self.motion_embedding = Embedding(obs_len)
            if 'mosa' in train_net or 'Layer' in train_net:
                self.encoder = YNetEncoderL(self.feature_channels, encoder_channels, train_net, position)
            else:
                self.encoder = YNetEncoderB(self.feature_channels, encoder_channels, train_net, position)
        else:
            raise ValueError('No network parameter is provided')
        self.goal_decoder = YNetDecoder(encoder_channels, decoder_channels, output_len=pred_len)
        self.traj_decoder = YNetDecoder(encoder_channels, decoder_channels, output_len=pred_len, traj=n_waypoints)
        self.softargmax_ = SoftArgmax2D(normalized_coordinates=False)
        self.encoder_channels = encoder_channels
        self._engine = None
        # the reference's scripts never choose an engine: YNET_BACKEND selects it for unchanged train.py / test.py runs
        self._backend = os.environ.get('YNET_BACKEND', 'fp32')
        if self._backend not in ('fp32', 'bf16', 'bf16x3'):
            raise ValueError(f"YNET_BACKEND={self._backend!r}: expected 'fp32', 'bf16x3' or 'bf16'")

    # ---- engine plumbing -------------------------------------------------------------------------
    def set_backend(self, backend):
        """'fp32' = CUDA-core reference-grade engine (<= 1e-3 parity); 'bf16' = tcgen05 tensor-core engine;
        'bf16x3' = split-bf16 tensor-core engine (<= 1e-3 parity, three MMAs per product)."""
        if backend not in ('fp32', 'bf16', 'bf16x3'):
            raise ValueError(f'unknown backend {backend!r}')
        object.__setattr__(self, '_backend', backend)
        object.__setattr__(self, '_engine', None)
        # (training graphs, autograd_engine: with 'bf16x3' the forward + data-gradient convs run on the tensor cores too)
        return self

    @property
    def engine(self):
        if self._engine is None:
            eng = {'bf16': YNetEngineTC, 'bf16x3': YNetEngineSplit}.get(self._backend, YNetEngine)(self)
            object.__setattr__(self, '_engine', eng)
        return self._engine

    def _training_graph(self, *inputs):
        """Differentiable executor (autograd_engine) when a parameter OR an input asks for gradients -- the latter is
        the saliency path of trainer.py:354-516 (frozen weights, ``scene_raw_img.requires_grad = True``)."""
        if not torch.is_grad_enabled():
            return False
        if any(p.requires_grad for p in self.parameters()):
            return True

        def walk(x):
            if isinstance(x, torch.Tensor):
                return x.requires_grad
            return isinstance(x, (tuple, list)) and any(walk(t) for t in x)
        return any(walk(x) for x in inputs)

    # ---- reference method set (ynet.py:551-600) ---------------------------------------------------
    def segmentation(self, image):
        return self.semantic_segmentation(image)

    def segmentation_cached(self, scene_id, image):
        """``segmentation(image)`` memoised per scene: evaluate() re-runs the frozen segmentation
        backbone for every scene in every round / epoch (evaluate.py:86-90, trainer.py:336-345); the semantic map only
        depends on the scene image and the backbone's weights, so it is kept until either changes."""
        ver = (image.data_ptr(), image._version, tuple(image.shape),
               tuple((q.data_ptr(), q._version) for q in self.semantic_segmentation.parameters()))
        cache = self.__dict__.setdefault('_semantic_cache', {})
        hit = cache.get(scene_id)
        if hit is not None and hit[0] == ver:
            return hit[1]
        out = self.semantic_segmentation(image)
        cache[scene_id] = (ver, out)
        return out

    def adapt_semantic(self, semantic_img):
        return semantic_img       # ynet.py:554-559: identity unless a semantic adapter exists (it cannot, see __init__)

    def pred_features(self, scene_map, motion_map):
        if self._training_graph(scene_map, motion_map):
            from .. import autograd_engine
            return autograd_engine.pred_features(self, scene_map, motion_map)
        return self.engine.pred_features(scene_map, motion_map)

    def pred_goal(self, features):
        if self._training_graph(features):
            from .. import autograd_engine
            return autograd_engine.decoder_logits(self, self.goal_decoder, 'goal_decoder', features)
        return self.engine.decoder_logits(self.goal_decoder, 'goal_decoder', features)

    def pred_traj(self, features):
        if self._training_graph(features):
            from .. import autograd_engine
            return autograd_engine.decoder_logits(self, self.traj_decoder, 'traj_decoder', features)
        return self.engine.decoder_logits(self.traj_decoder, 'traj_decoder', features)

    def pred_traj_softargmax(self, features):
        """pred_traj followed by softargmax (evaluate.py:263-264) without materialising the logits."""
        return self.engine.decoder_softargmax(self.traj_decoder, 'traj_decoder', features)

    def softmax(self, x):
        return ops.spatial_softmax(x)

    def softargmax(self, output):
        return self.softargmax_(output)

    def sigmoid(self, output):
        B, C = output.shape[:2]
        return ops.sigmoid_select(output, list(range(C)), 1.0)

    def softargmax_on_softmax_map(self, x):
        return ops.expectation2d(x)

    def forward(self, *a, **k):  # the reference defines no forward either
        raise NotImplementedError('use pred_features / pred_goal / pred_traj')


__all__ = ['YNet', 'YNetEncoder', 'YNetEncoderL', 'YNetEncoderB', 'YNetEncoderFusion', 'YNetDecoder',
           'get_conv2d', 'ChannelCat']
