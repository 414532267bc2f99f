"""Drop-in mirror of ``hcat.unet.Unet_Constructor`` (`hcat/unet.py:15-196`) on the B200 kernels.

Same constructor kwargs, ``model_specification``, ``forward``, ``save``/``load`` and a ``state_dict``
with the reference's keys and tensor shapes, so checkpoints interchange both ways.  The submodules
(``down_steps.i.conv1`` ... ``up_steps.i.up_conv``, ``out_conv``) are ordinary ``torch.nn`` modules used
ONLY as parameter containers -- they give the reference's registration order, initialisation (same RNG
consumption, so ``torch.manual_seed(s); Unet_Constructor(...)`` yields the reference's weights) and
``.to()/.parameters()/.train()/.eval()`` semantics.  Their ``forward`` is never called: every arithmetic
step of ``forward``/``backward`` is a hand-written sm_100a kernel of ``libhcunet_b200.so`` reached through
the C ABI (``include/hcunet_b200.h``).  There is no CPU or PyTorch fallback: a non-CUDA input, or a
missing library, raises.

Deviations from the reference, all deliberate and documented in DESIGN.md:
  * ``image_dimensions=2`` constructs (the reference raises ``RuntimeError('fuck', ...)`` at
    `unet.py:293-303`); parity for it is checked against the reference with ConvTranspose3d aliased during
    construction.
  * ``precision``: 'fp32' (default; FFMA kernels, fp32 activations, rel-L2 <= 1e-5 vs the reference) or
    'mixed' (fp16 activations, fp32 accumulate on tcgen05 tensor cores, rel-L2 <= 2e-3).  Select with
    ``model.precision = 'mixed'`` or the ``HCUNET_PRECISION`` environment variable.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List

import torch
import torch.nn as nn

from .engine import UnetEngine

_PRECISIONS = ("fp32", "mixed")


class Down(nn.Module):
    """Parameter container mirroring `unet.py:236-266` (conv1, conv2, batch1, batch2)."""

    def __init__(self, conv_functions, in_channels, out_channels, kernel, dilation, groups):
        super().__init__()
        self.conv1 = conv_functions[0](in_channels, out_channels, kernel["conv1"], dilation=dilation["conv1"],
                                       groups=groups["conv1"], padding=0)
        self.conv2 = conv_functions[0](out_channels, out_channels, kernel["conv2"], dilation=dilation["conv2"],
                                       groups=groups["conv2"], padding=0)
        self.batch1 = conv_functions[3](out_channels)
        self.batch2 = conv_functions[3](out_channels)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("hcunet_b200.Down is a parameter container; call the Unet_Constructor")


class Up(nn.Module):
    """Parameter container mirroring `unet.py:269-315` (conv1, conv2, up_conv, batch1, batch2).

    ``conv1`` takes ``in_channels`` (= 2 * out_channels) inputs because the reference feeds it
    ``cat(x_up, crop(x_up, skip))`` (`unet.py:311-312`)."""

    def __init__(self, conv_functions, in_channels, out_channels, kernel, upsample_kernel, upsample_stride, dilation,
                 groups):
        super().__init__()
        self.conv1 = conv_functions[0](in_channels, out_channels, kernel["conv1"], dilation=dilation["conv1"],
                                       groups=groups["conv1"], padding=0)
        self.conv2 = conv_functions[0](out_channels, out_channels, kernel["conv2"], dilation=dilation["conv2"],
                                       groups=groups["conv2"], padding=0)
        self.up_conv = conv_functions[1](in_channels, out_channels, upsample_kernel, stride=upsample_stride, padding=0)
        self.lin_up = False
        self.batch1 = conv_functions[3](out_channels)
        self.batch2 = conv_functions[3](out_channels)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("hcunet_b200.Up is a parameter container; call the Unet_Constructor")


class _UnetFunction(torch.autograd.Function):
    """forward / backward of the whole U-Net as one autograd node over the C ABI."""

    @staticmethod
    def forward(ctx, module, need_grad, x, *params):
        names = module._param_names
        pd = {n: p.detach() for n, p in zip(names, params)}
        logits, state = module._engine.forward(pd, module._buffer_dict(), x.detach(), training=module.training,
                                               save=need_grad, precision=module.precision, prelaid=module._x_prelaid)
        ctx.module = module
        ctx.state = state
        ctx.pd = pd
        ctx.x_needs_grad = x.requires_grad
        ctx.set_materialize_grads(False)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        if dlogits is None:
            return (None, None, None) + (None,) * len(module._param_names)
        if ctx.state is None or ctx.state[1] is None:
            raise RuntimeError("hcunet_b200: backward through a forward that saved no state")
        grads, dx = module._engine.backward(ctx.pd, ctx.state, dlogits, need_dx=ctx.x_needs_grad)
        ctx.state = None
        out = tuple(grads.get(n) for n in module._param_names)
        return (None, None, dx) + out


class Unet_Constructor(nn.Module):
    def __init__(self,
                 image_dimensions=2,
                 in_channels=3,
                 out_channels=2,
                 feature_sizes=[32, 64, 128, 256, 512, 1024],
                 kernel=(3, 3),
                 upsample_kernel=(2, 2),
                 max_pool_kernel=(2, 2),
                 upsample_stride=2,
                 dilation=1,
                 groups=1,
                 ):
        super(Unet_Constructor, self).__init__()
        if image_dimensions == 2:
            conv_functions = (nn.Conv2d, nn.ConvTranspose2d, nn.MaxPool2d, nn.BatchNorm2d)
        elif image_dimensions == 3:
            conv_functions = (nn.Conv3d, nn.ConvTranspose3d, nn.MaxPool3d, nn.BatchNorm3d)
        else:
            raise ValueError(f'Does not support {image_dimensions} dimensional images')  # unet.py:53

        # unet.py:59-64
        if type(kernel) is tuple:
            kernel = {'conv1': kernel, 'conv2': kernel}
        if type(dilation) is int or type(dilation) is tuple:
            dilation = {'conv1': dilation, 'conv2': dilation}
        if type(groups) is int or type(groups) is tuple:
            groups = {'conv1': groups, 'conv2': groups}

        # unet.py:67-71
        if len(feature_sizes) < 2:
            raise ValueError(f'The Number of Features must be at least 2, not {len(feature_sizes)}')
        for i, f in enumerate(feature_sizes[0:-1:1]):
            assert f * 2 == feature_sizes[i + 1], \
                f'Feature Sizes must be multiples of two from each other: {f} != {feature_sizes[i - 1]}*2'

        self.model_specification = {  # unet.py:74-85
            'image_dimensions': image_dimensions,
            'in_channels': in_channels,
            'out_channels': out_channels,
            'feature_sizes': feature_sizes,
            'kernel': kernel,
            'upsample_kernel': upsample_kernel,
            'max_pool_kernel': max_pool_kernel,
            'upsample_stride': upsample_stride,
            'dilation': dilation,
            'groups': groups
        }

        # construction (and therefore RNG) order of unet.py:87-123: downs, ups, out_conv
        down_steps: List[nn.Module] = [Down(conv_functions, in_channels, feature_sizes[0], kernel, dilation, groups)]
        for i in range(1, len(feature_sizes)):
            down_steps.append(Down(conv_functions, feature_sizes[i - 1], feature_sizes[i], kernel, dilation, groups))
        up_steps: List[nn.Module] = []
        i = -2
        for f in feature_sizes[:0:-1]:
            up_steps.append(Up(conv_functions, f, feature_sizes[i], kernel, upsample_kernel, upsample_stride, dilation,
                               groups))
            i += -1
        # registration order of the reference: out_conv first, then the two ModuleLists (unet.py:120-122)
        self.out_conv = conv_functions[0](feature_sizes[0], out_channels, 1)
        self.down_steps = nn.ModuleList(down_steps)
        self.up_steps = nn.ModuleList(up_steps)
        self.max_pool = conv_functions[2](max_pool_kernel)

        self.precision = os.environ.get("HCUNET_PRECISION", "fp32")
        if self.precision not in _PRECISIONS:
            raise ValueError(f"HCUNET_PRECISION must be one of {_PRECISIONS}, not {self.precision}")
        self._engine = UnetEngine(self.model_specification)
        self._x_prelaid = False
        self._param_names = [n for n, _ in self.named_parameters()]
        self._buffer_names = [n for n, _ in self.named_buffers()]

    # ---- plumbing ---------------------------------------------------------------------------
    def _buffer_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_buffers())

    def forward(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"expected a torch.Tensor, got {type(x)}")
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {_PRECISIONS}, not {self.precision}")
        if not x.is_cuda:
            raise RuntimeError("hcunet_b200.Unet_Constructor runs on CUDA (sm_100a) only: there is no CPU fallback. "
                               "Move the model and the input to a B200 (`.cuda()`).")
        params = [p for _, p in self.named_parameters()]
        if params[0].device != x.device:
            raise RuntimeError(f"Input type ({x.device}) and weight type ({params[0].device}) should be the same")
        # geometry errors (too-small inputs, channel mismatch) are raised by the planner with the
        # reference's RuntimeError semantics before any kernel is launched
        self._engine.plan(tuple(x.shape))
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        self._x_prelaid = bool(getattr(x, "_hcu_cl8", False))   # hcunet_b200.loader.StackLoader.image output
        return _UnetFunction.apply(self, need_grad, x, *params)

    # ---- unet.py:145-165 --------------------------------------------------------------------
    def save(self, filename, hyperparameters=None):
        model = {'state_dict': self.state_dict(),
                 'model_specifications': self.model_specification,
                 'hyperparameters': hyperparameters}
        python_files = {}
        python_files_list = glob.glob('./**/*.py', recursive=True)
        for f in glob.glob('./**/*.ipynb', recursive=True):
            python_files_list.append(f)
        for f in python_files_list:
            with open(f, 'r') as file:
                python_files[f] = file.read()
        model['python_files'] = python_files
        model['tree_structure'] = glob.glob('**/*', recursive=True)
        torch.save(model, filename)
        return None

    # ---- unet.py:167-196 --------------------------------------------------------------------
    def load(self, filename, to_cuda=True):
        if torch.cuda.is_available() and to_cuda:
            device = 'cuda:0'
        else:
            device = 'cpu'
        model = torch.load(filename, map_location=device, weights_only=False)
        model_specification = model['model_specifications']
        precision = getattr(self, "precision", None)
        self.__init__(
            image_dimensions=model_specification['image_dimensions'],
            in_channels=model_specification['in_channels'],
            out_channels=model_specification['out_channels'],
            feature_sizes=model_specification['feature_sizes'],
            kernel=model_specification['kernel'],
            upsample_kernel=model_specification['upsample_kernel'],
            max_pool_kernel=model_specification['max_pool_kernel'],
            upsample_stride=model_specification['upsample_stride'],
            dilation=model_specification['dilation'],
            groups=model_specification['groups'],
        )
        if precision is not None:
            self.precision = precision
        self.load_state_dict(model['state_dict'])
        self.eval()
        try:
            return model['hyperparameters']
        except KeyError:
            return None

    def evaluate(self, image: torch.Tensor):
        """`unet.py:198-233`, as unfinished as in the reference: same argument checks (`ValueError` for a non-tensor,
        `ImportError` for a channel mismatch -- sic), `eval()`, reflection padding by (100, 100, 8), a forward pass over every
        200 x 200 XY slice of the padded image whose result is dropped, and `None`.  The tiled inference callers really use
        is `hcunet_b200.segment.predict_segmentation_mask` (`segment.py:21-136`)."""
        if not isinstance(image, torch.Tensor):
            raise ValueError(f'Expected image type of torch.Tensor, not {type(image)}')
        if image.shape[1] != self.model_specification['in_channels']:
            raise ImportError(f'Image expected to have {self.model_specification["in_channels"]} not {image.shape[1]}')
        from .segment import pad_image_with_reflections

        self.eval()
        pad = (100, 100, 8)
        skip = 200
        device = next(self.parameters()).device
        padded_image = pad_image_with_reflections(image.to(device), pad).float()
        with torch.no_grad():
            for x in range(0, padded_image.shape[2], skip):
                for y in range(0, padded_image.shape[3], skip):
                    slice_to_eval = padded_image[:, :, x:x + skip, y:y + skip, :]
                    self.forward(slice_to_eval)      # the reference slices the result into `mask_slice` and drops it
        return None
