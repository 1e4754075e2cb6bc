"""Plain-torch glue networks of the full camera+LiDAR BEVFusion frame (BASELINE config C4), shapes per
SURVEY.md Appendix E.  NOT part of the product: random frozen weights, cuDNN / cuBLAS kernels, used by
`bench.py --workload c4` only, to measure the hot path (bev_pool, voxelize, SparseEncoder: this repo's
kernels) inside the frame the reference's FPS protocol times (tools/benchmark.py:56-85,
mmdet3d/models/fusion_models/bevfusion.py:274-388).

    img [B,6,3,256,704] -> SwinT-T -> GeneralizedLSSFPN -> [B*6,256,32,88]
    points -> per-camera depth image (this repo's depth rasteriser) -> dtransform -> depthnet
           -> softmax depth [B,6,118,32,88] , context [B,6,32,88,80] -> fused lift (x) bev_pool -> [B,80,360,360]
           -> downsample -> [B,80,180,180]
    points -> hard voxelize + mean -> SparseEncoder -> [B,256,180,180]
    ConvFuser -> SECOND -> SECONDFPN -> TransFusionHead (200 proposals, 1 decoder layer) -> boxes"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def conv_bn_relu(cin, cout, k, s=1, p=0, eps=1e-5, bias=False):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p, bias=bias), nn.BatchNorm2d(cout, eps=eps), nn.ReLU(True))


# ---- SwinTransformer-T (mmdet 2.20 layout: embed 96, depths 2-2-6-2, heads 3-6-12-24, window 7) ----------
class WindowAttention(nn.Module):
    def __init__(self, dim, heads, window):
        super().__init__()
        self.dim, self.heads, self.window = dim, heads, window
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        self.rel_bias = nn.Parameter(torch.zeros((2 * window - 1) ** 2, heads))
        coords = torch.stack(torch.meshgrid(torch.arange(window), torch.arange(window), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0) + window - 1
        self.register_buffer("rel_index", (rel[..., 0] * (2 * window - 1) + rel[..., 1]).long(), persistent=False)

    def forward(self, x, mask):                       # x [nW*B, N, C]
        Bw, N, C = x.shape
        qkv = self.qkv(x).view(Bw, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        bias = self.rel_bias[self.rel_index.view(-1)].view(N, N, self.heads).permute(2, 0, 1).unsqueeze(0)
        if mask is not None:                          # [nW, N, N]
            nW = mask.shape[0]
            bias = (bias + mask.unsqueeze(1)).unsqueeze(0).expand(Bw // nW, -1, -1, -1, -1).reshape(Bw, self.heads, N, N)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=bias.to(x.dtype))
        return self.proj(out.transpose(1, 2).reshape(Bw, N, C))


class SwinBlock(nn.Module):
    def __init__(self, dim, heads, window, shift):
        super().__init__()
        self.window, self.shift = window, shift
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, heads, window)
        self.mlp = nn.Sequential(nn.Linear(dim, 4 * dim), nn.GELU(), nn.Linear(4 * dim, dim))

    def forward(self, x, H, W):                       # x [B, H*W, C]
        B, _, C = x.shape
        w = self.window
        h = self.norm1(x).view(B, H, W, C)
        ph, pw = (w - H % w) % w, (w - W % w) % w
        h = F.pad(h, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        mask = None
        if self.shift:
            h = torch.roll(h, (-self.shift, -self.shift), (1, 2))
            img = torch.zeros((1, Hp, Wp, 1), device=x.device)
            cnt = 0
            for hs in (slice(0, -w), slice(-w, -self.shift), slice(-self.shift, None)):
                for ws in (slice(0, -w), slice(-w, -self.shift), slice(-self.shift, None)):
                    img[:, hs, ws] = cnt
                    cnt += 1
            mw = img.view(1, Hp // w, w, Wp // w, w, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, w * w)
            mask = mw[:, None, :] - mw[:, :, None]
            mask = torch.where(mask != 0, torch.full_like(mask, -100.0), torch.zeros_like(mask))
        win = h.view(B, Hp // w, w, Wp // w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, w * w, C)
        win = self.attn(win, mask)
        h = win.view(B, Hp // w, Wp // w, w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        if self.shift:
            h = torch.roll(h, (self.shift, self.shift), (1, 2))
        x = x + h[:, :H, :W].reshape(B, H * W, C)
        return x + self.mlp(self.norm2(x))


class SwinT(nn.Module):
    def __init__(self, embed=96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), window=7, out_indices=(1, 2, 3)):
        super().__init__()
        self.patch = nn.Conv2d(3, embed, 4, 4)
        self.patch_norm = nn.LayerNorm(embed)
        self.stages, self.merges, self.out_norms = nn.ModuleList(), nn.ModuleList(), nn.ModuleDict()
        dim = embed
        self.out_indices = out_indices
        for i, (d, hd) in enumerate(zip(depths, heads)):
            self.stages.append(nn.ModuleList([SwinBlock(dim, hd, window, 0 if j % 2 == 0 else window // 2) for j in range(d)]))
            if i in out_indices:
                self.out_norms[str(i)] = nn.LayerNorm(dim)
            if i + 1 < len(depths):
                self.merges.append(nn.Sequential(nn.LayerNorm(4 * dim), nn.Linear(4 * dim, 2 * dim, bias=False)))
                dim *= 2

    def forward(self, img):
        x = self.patch(img)
        B, C, H, W = x.shape
        x = self.patch_norm(x.flatten(2).transpose(1, 2))
        outs = []
        for i, blocks in enumerate(self.stages):
            for blk in blocks:
                x = blk(x, H, W)
            if i in self.out_indices:
                outs.append(self.out_norms[str(i)](x).view(B, H, W, -1).permute(0, 3, 1, 2).contiguous())
            if i < len(self.merges):
                x = x.view(B, H, W, -1)
                x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
                x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
                H, W = (H + 1) // 2, (W + 1) // 2
                x = self.merges[i](x.view(B, H * W, -1))
        return outs


class GeneralizedLSSFPN(nn.Module):
    """necks/generalized_lss.py:48-103: top-down, upsample -> cat -> 1x1 ConvModule -> 3x3 ConvModule."""

    def __init__(self, in_channels=(192, 384, 768), out_channels=256):
        super().__init__()
        self.lateral, self.fpn = nn.ModuleList(), nn.ModuleList()
        for i in range(len(in_channels) - 1):
            cin = in_channels[i] + (in_channels[i + 1] if i == len(in_channels) - 2 else out_channels)
            self.lateral.append(conv_bn_relu(cin, out_channels, 1))
            self.fpn.append(conv_bn_relu(out_channels, out_channels, 3, p=1))

    def forward(self, feats):
        lat = list(feats)
        for i in range(len(feats) - 2, -1, -1):
            up = F.interpolate(lat[i + 1], size=lat[i].shape[2:], mode="bilinear", align_corners=False)
            lat[i] = self.fpn[i](self.lateral[i](torch.cat([lat[i], up], 1)))
        return lat[0]


class DepthLSSNets(nn.Module):
    """vtransforms/depth_lss.py:38-79: dtransform, depthnet, BEV downsample."""

    def __init__(self, in_channels=256, C=80, D=118):
        super().__init__()
        self.C, self.D = C, D
        self.dtransform = nn.Sequential(conv_bn_relu(1, 8, 1, bias=True), conv_bn_relu(8, 32, 5, 4, 2, bias=True),
                                        conv_bn_relu(32, 64, 5, 2, 2, bias=True))
        self.depthnet = nn.Sequential(conv_bn_relu(in_channels + 64, in_channels, 3, p=1, bias=True),
                                      conv_bn_relu(in_channels, in_channels, 3, p=1, bias=True),
                                      nn.Conv2d(in_channels, D + C, 1))
        self.downsample = nn.Sequential(conv_bn_relu(C, C, 3, p=1), conv_bn_relu(C, C, 3, 2, 1), conv_bn_relu(C, C, 3, p=1))

    def lift_inputs(self, x, d):
        """x [B*N,256,fH,fW], d [B*N,1,H,W] -> depth [B*N,D,fH,fW] (softmax), context [B*N,fH,fW,C]."""
        y = self.depthnet(torch.cat([self.dtransform(d), x], 1))
        depth = y[:, :self.D].softmax(dim=1)
        ctx = y[:, self.D:self.D + self.C].permute(0, 2, 3, 1).contiguous()
        return depth.contiguous(), ctx


class ConvFuser(nn.Sequential):
    def __init__(self, in_channels=(80, 256), out_channels=256):
        super().__init__(nn.Conv2d(sum(in_channels), out_channels, 3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
                         nn.ReLU(True))


class SECOND(nn.Module):
    def __init__(self, cin=256, outs=(128, 256), layer_nums=(5, 5), strides=(1, 2)):
        super().__init__()
        self.blocks = nn.ModuleList()
        for co, n, s in zip(outs, layer_nums, strides):
            layers = [conv_bn_relu(cin, co, 3, s, 1, eps=1e-3)] + [conv_bn_relu(co, co, 3, 1, 1, eps=1e-3) for _ in range(n)]
            self.blocks.append(nn.Sequential(*layers))
            cin = co

    def forward(self, x):
        outs = []
        for b in self.blocks:
            x = b(x)
            outs.append(x)
        return outs


class SECONDFPN(nn.Module):
    def __init__(self, ins=(128, 256), outs=(256, 256)):
        super().__init__()
        self.up0 = nn.Sequential(nn.Conv2d(ins[0], outs[0], 1, bias=False), nn.BatchNorm2d(outs[0], eps=1e-3), nn.ReLU(True))
        self.up1 = nn.Sequential(nn.ConvTranspose2d(ins[1], outs[1], 2, 2, bias=False), nn.BatchNorm2d(outs[1], eps=1e-3),
                                 nn.ReLU(True))

    def forward(self, xs):
        return torch.cat([self.up0(xs[0]), self.up1(xs[1])], 1)


class PredFFN(nn.Module):
    def __init__(self, cin=128, heads=(("center", 2), ("height", 1), ("dim", 3), ("rot", 2), ("vel", 2), ("heatmap", 10))):
        super().__init__()
        self.heads = nn.ModuleDict({k: nn.Sequential(nn.Conv1d(cin, 64, 1), nn.BatchNorm1d(64), nn.ReLU(True), nn.Conv1d(64, n, 1))
                                    for k, n in heads})

    def forward(self, x):
        return {k: h(x) for k, h in self.heads.items()}


class TransFusionHead(nn.Module):
    """heads/bbox/transfusion.py:224-275 + utils/transformer.py:14-112, inference path only."""

    def __init__(self, cin=512, hidden=128, proposals=200, classes=10, heads=8, ffn=256):
        super().__init__()
        self.P, self.K = proposals, classes
        self.shared_conv = nn.Conv2d(cin, hidden, 3, padding=1)
        self.heatmap_head = nn.Sequential(conv_bn_relu(hidden, hidden, 3, p=1, bias=True), nn.Conv2d(hidden, classes, 3, padding=1))
        self.class_encoding = nn.Conv1d(classes, hidden, 1)
        self.query_pos = nn.Sequential(nn.Conv1d(2, hidden, 1), nn.BatchNorm1d(hidden), nn.ReLU(True), nn.Conv1d(hidden, hidden, 1))
        self.key_pos = nn.Sequential(nn.Conv1d(2, hidden, 1), nn.BatchNorm1d(hidden), nn.ReLU(True), nn.Conv1d(hidden, hidden, 1))
        self.self_attn = nn.MultiheadAttention(hidden, heads, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(hidden, heads, batch_first=True)
        self.ffn = nn.Sequential(nn.Linear(hidden, ffn), nn.ReLU(True), nn.Linear(ffn, hidden))
        self.n1, self.n2, self.n3 = nn.LayerNorm(hidden), nn.LayerNorm(hidden), nn.LayerNorm(hidden)
        self.pred = PredFFN(hidden)

    def forward(self, x):
        B = x.shape[0]
        feat = self.shared_conv(x)
        H, W = feat.shape[2:]
        flat = feat.flatten(2)                                         # [B, 128, HW]
        ys, xs = torch.meshgrid(torch.arange(H, device=x.device), torch.arange(W, device=x.device), indexing="ij")
        bev_pos = torch.stack([xs + 0.5, ys + 0.5], 0).float().flatten(1).unsqueeze(0).expand(B, -1, -1)   # [B,2,HW]
        heat = self.heatmap_head(feat).sigmoid()
        local_max = F.max_pool2d(heat, 3, 1, 1)
        local_max[:, 8:] = heat[:, 8:]                                 # pedestrian / cone: no suppression
        heat = (heat * (heat == local_max)).flatten(1)                 # [B, K*HW]
        score, top = heat.topk(self.P, dim=1)
        cls, pos = top // (H * W), top % (H * W)
        q = flat.gather(2, pos.unsqueeze(1).expand(-1, flat.shape[1], -1))            # [B,128,P]
        q = q + self.class_encoding(F.one_hot(cls, self.K).permute(0, 2, 1).float())
        qpos = bev_pos.gather(2, pos.unsqueeze(1).expand(-1, 2, -1))
        qe, ke = self.query_pos(qpos), self.key_pos(bev_pos)
        qt, kt = q.transpose(1, 2), flat.transpose(1, 2)
        qpe, kpe = qe.transpose(1, 2), ke.transpose(1, 2)
        qt = self.n1(qt + self.self_attn(qt + qpe, qt + qpe, qt)[0])
        qt = self.n2(qt + self.cross_attn(qt + qpe, kt + kpe, kt)[0])
        qt = self.n3(qt + self.ffn(qt))
        out = self.pred(qt.transpose(1, 2))
        out["center"] = out["center"] + qpos
        # TransFusionBBoxCoder.decode (transfusion_bbox_coder.py:60-123)
        scores = out["heatmap"].sigmoid() * score.unsqueeze(1) * F.one_hot(cls, self.K).permute(0, 2, 1)
        final_score, labels = scores.max(1)
        cx = out["center"][:, 0] * 8 * 0.075 - 54.0
        cy = out["center"][:, 1] * 8 * 0.075 - 54.0
        rot = torch.atan2(out["rot"][:, 0], out["rot"][:, 1])
        boxes = torch.stack([cx, cy, out["height"][:, 0], *out["dim"].exp().unbind(1), rot, *out["vel"].unbind(1)], -1)
        return boxes, final_score, labels


class GlueNets(nn.Module):
    """Everything of the C4 frame that is not the hot path."""

    def __init__(self):
        super().__init__()
        self.backbone, self.neck, self.lss = SwinT(), GeneralizedLSSFPN(), DepthLSSNets()
        self.fuser, self.second, self.secondfpn, self.head = ConvFuser(), SECOND(), SECONDFPN(), TransFusionHead()

    def camera_features(self, img):                   # [B,6,3,256,704] -> [B*6,256,32,88]
        B, N = img.shape[:2]
        return self.neck(self.backbone(img.flatten(0, 1)))

    def decode(self, cam_bev, lidar_bev_in_buf):
        """`lidar_bev_in_buf` = the fuser input [B, 80+256, 180, 180] whose LiDAR channels the encoder wrote in place
        and whose camera channels get the downsampled camera BEV here."""
        lidar_bev_in_buf[:, :80] = self.lss.downsample(cam_bev)
        x = self.fuser(lidar_bev_in_buf)
        return self.head(self.secondfpn(self.second(x)))
