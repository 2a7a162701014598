"""Mirror of the pixelSplat decoder that drives the rasterizer in GGRt
(/root/reference/ggrt/model/pixelsplat/decoder/decoder_splatting_cuda.py:29-85,
decoder.py:20-23, ../types.py:7-12): flattens (batch, view), renders colour and, when a
depth mode is given, a second degree-0 pass for depth."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from .render import (DepthRenderingMode, render_color_and_depth_cuda, render_cuda, render_depth_cuda,
                     render_views_device, render_views_fast)


@dataclass
class Gaussians:
    means: Tensor  # [b,g,3]
    covariances: Tensor  # [b,g,3,3]
    harmonics: Tensor  # [b,g,3,d_sh]
    opacities: Tensor  # [b,g]


@dataclass
class DecoderOutput:
    color: Tensor  # [b,v,3,h,w]
    depth: Optional[Tensor]  # [b,v,h,w]


def _per_view(t: Tensor, v: int) -> Tensor:
    """[b, ...] -> [(b v), ...] without copying until the rasterizer needs contiguity."""
    return t[:, None].expand(t.shape[0], v, *t.shape[1:]).reshape(t.shape[0] * v, *t.shape[1:])


class DecoderSplattingCUDA(nn.Module):
    def __init__(self, background_color=(0.0, 0.0, 0.0), fused_depth: bool = False, fast_glue: bool = False,
                 device_glue: bool = False, view_streams: int = 1) -> None:
        """fused_depth: render colour and depth in one rasterization (same outputs as the reference's two
        passes; roughly half the rasterizer work when a depth mode is requested).
        fast_glue: additionally skip the reference glue's per-view replication / rescale / gather / permute
        copies and its per-view host syncs (render_views_fast); implies fused_depth.
        device_glue: additionally set up all cameras in one kernel and keep them on the device -- no host
        synchronisation at all, the depth channel evaluated inside the rasterizer (render_views_device; depth modes
        other than "depth" fall back to fast_glue); implies fast_glue.
        view_streams (with device_glue): issue the (batch x view) rasterizations round-robin on this many CUDA streams so
        that independent views overlap on the GPU (render_views_device)."""
        super().__init__()
        self.background_color = torch.tensor(background_color, dtype=torch.float32)
        self.fused_depth = fused_depth or fast_glue or device_glue
        self.fast_glue = fast_glue or device_glue
        self.device_glue = device_glue
        self.view_streams = int(view_streams)
        self._bg = {}

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple, depth_mode: Optional[DepthRenderingMode] = None) -> DecoderOutput:
        b, v = extrinsics.shape[:2]
        key = (far.device, b * v)
        if key not in self._bg:  # cached: one host-to-device copy per (device, view count), not per call
            self._bg[key] = self.background_color.to(far.device)[None].expand(b * v, 3).contiguous()
        bg = self._bg[key]
        if self.device_glue and depth_mode in (None, "depth"):
            color, depth = render_views_device(
                extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(), image_shape, bg,
                gaussians.means, gaussians.covariances, gaussians.harmonics, gaussians.opacities,
                view_to_scene=[i // v for i in range(b * v)], depth=depth_mode is not None, streams=self.view_streams)
            return DecoderOutput(color.unflatten(0, (b, v)), None if depth is None else depth.unflatten(0, (b, v)))
        if self.fast_glue:
            color, depth = render_views_fast(
                extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(), image_shape, bg,
                gaussians.means, gaussians.covariances, gaussians.harmonics, gaussians.opacities,
                view_to_scene=[i // v for i in range(b * v)], depth_mode=depth_mode)
            return DecoderOutput(color.unflatten(0, (b, v)), None if depth is None else depth.unflatten(0, (b, v)))
        if self.fused_depth and depth_mode is not None:
            color, depth = render_color_and_depth_cuda(
                extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(), image_shape, bg,
                _per_view(gaussians.means, v), _per_view(gaussians.covariances, v), _per_view(gaussians.harmonics, v),
                _per_view(gaussians.opacities, v), mode=depth_mode)
            return DecoderOutput(color.unflatten(0, (b, v)), depth.unflatten(0, (b, v)))
        color = render_cuda(extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(),
                            image_shape, bg, _per_view(gaussians.means, v), _per_view(gaussians.covariances, v),
                            _per_view(gaussians.harmonics, v), _per_view(gaussians.opacities, v))
        color = color.unflatten(0, (b, v))
        depth = None
        if depth_mode is not None:
            depth = self.render_depth(gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode)
        return DecoderOutput(color, depth)

    def capture(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple, grad_color: Tensor, grad_depth: Optional[Tensor] = None, streams: int = 2):
        """A training loop that renders the same shapes every step: all v target views of ONE scene (b = 1) -- forward,
        depth channel and backward -- frozen into one CUDA-graph launch (graph.CapturedViews).  `grad_color` [v,3,h,w] /
        `grad_depth` [v,h,w] are the buffers the loss writes dL/dimage into; the Gaussians and cameras are read from the
        given tensors at every replay."""
        from .graph import CapturedViews

        b, v = extrinsics.shape[:2]
        if b != 1:
            raise ValueError("capture() takes one scene (b = 1); build one CapturedViews per scene")
        bg = self.background_color.to(far.device)[None].expand(v, 3).contiguous()
        return CapturedViews(extrinsics[0], intrinsics[0], near[0], far[0], image_shape, bg, gaussians.means[0],
                             gaussians.covariances[0], gaussians.harmonics[0], gaussians.opacities[0],
                             grad_color=grad_color, grad_depth=grad_depth, streams=streams)

    def render_depth(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                     image_shape: tuple, mode: DepthRenderingMode = "depth") -> Tensor:
        b, v = extrinsics.shape[:2]
        result = render_depth_cuda(extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(),
                                   image_shape, _per_view(gaussians.means, v), _per_view(gaussians.covariances, v),
                                   _per_view(gaussians.opacities, v), mode=mode)
        return result.unflatten(0, (b, v))
