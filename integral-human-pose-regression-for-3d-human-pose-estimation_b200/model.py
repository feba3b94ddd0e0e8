"""Model-side mirror of /root/reference/main/model.py for the end-to-end harness.

Same module tree and parameter names as the reference (`backbone.*`, `head.deconv_layers.{0,1,3,4,6,7}.*`,
`head.final_layer.{weight,bias}`) so checkpoints written by the reference (with their DataParallel `module.` prefix,
see trainer.load_reference_checkpoint) load unchanged.  `ResPoseNet.forward(input_img, target=None)` is the superset
contract `north_star` asks for: without a target it returns the heat-map exactly like the reference
(model.py:99-103); with `target = {'coord', 'vis', 'have_depth'}` it returns the integral L1 loss computed by the
sm_100a path (what main/train.py:64-67 does in two calls).
"""
import torch
import torch.nn as nn

from .functional import DeferredHeatmap, deconv_bn_relu, deconv_bn_relu_train, flip_merge, fused_head_integral_l1_loss, fused_head_soft_argmax
from .nets.loss import JointLocationLoss, soft_argmax
from .nets.resnet import ResNetBackbone


class HeadNet(nn.Module):
    """3 x (ConvTranspose2d k4 s2 p1 -> BatchNorm -> ReLU) then the 1x1 conv to J*depth_dim channels (model.py:5-44)."""

    def __init__(self, joint_num, depth_dim=64, inplanes=2048, outplanes=256):
        super().__init__()
        layers = []
        for _ in range(3):
            layers += [nn.ConvTranspose2d(inplanes, outplanes, kernel_size=4, stride=2, padding=1, output_padding=0, bias=False),
                       nn.BatchNorm2d(outplanes), nn.ReLU(inplace=True)]
            inplanes = outplanes
        self.deconv_layers = nn.Sequential(*layers)
        self.final_layer = nn.Conv2d(inplanes, joint_num * depth_dim, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        return self.final_layer(self.deconv_layers(x))

    def features(self, x, fused_training=False):
        """deconv_layers(x).  At inference (eval mode, no autograd) every ConvTranspose2d(·, 256) + BatchNorm + ReLU block whose input
        map is 16 or 32 wide -- the second and the third block for the reference's 256 x 256 input -- runs as the tensor-core kernel
        K9 and the last one lands in the bf16 channels_last layout K3 reads.  With ``fused_training=True`` the
        same blocks (256 -> 256 channels) run in TRAINING mode through ``deconv_bn_relu_train`` (K9's training forward with batch statistics
        from the epilogue, K10 BatchNorm / ReLU passes, K9's input-gradient GEMM).  Everything else, and every other case, is the stock
        module stack."""
        dl = self.deconv_layers
        infer = not self.training and not torch.is_grad_enabled()
        train = bool(fused_training) and self.training and torch.is_grad_enabled()
        if not x.is_cuda or not (infer or train):
            return dl(x)
        for i in range(0, len(dl), 3):
            conv, bn = dl[i], dl[i + 1]
            h, w = x.shape[2], x.shape[3]
            ok = (isinstance(conv, nn.ConvTranspose2d) and isinstance(bn, nn.BatchNorm2d) and conv.out_channels == 256 and conv.in_channels % 64 == 0
                  and conv.in_channels <= 1024 and w in (16, 32) and h % (256 // w) == 0 and bn.affine)
            if ok and infer and bn.track_running_stats and not bn.training:
                x = deconv_bn_relu(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            elif ok and train and bn.training and conv.in_channels == 256 and x.shape[0] > 0 and bn.momentum is not None:
                # (a BatchNorm that is frozen -- .eval() inside a training head -- uses running statistics WITH gradients: the stock modules)
                x = deconv_bn_relu_train(x, conv.weight, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                                         bn.running_var if bn.track_running_stats else None, bn.momentum, bn.eps)
                if bn.track_running_stats and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
            else:
                x = dl[i:i + 3](x)
        return x

    def init_weights(self):                            # model.py:46-56
        for m in self.modules():
            if isinstance(m, (nn.ConvTranspose2d, nn.Conv2d)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class ResPoseNet(nn.Module):
    def __init__(self, backbone, head, joint_num=None, fused_head=False, deferred=False, fused_deconv=None):
        """fused_head=True: final_layer + soft-argmax (+ loss) run as the tensor-core kernels K3 / K4 and the
        (B, J*D, H, W) heat-map is never stored (same parameters, same checkpoints).
        deferred=True (with fused_head): ``forward(img)`` without a target returns a ``DeferredHeatmap`` instead of the tensor, so
        the reference's two-call sequences (train.py:64-67, test.py:62-65) reach K3 / K4 without being rewritten."""
        super().__init__()
        self.backbone = backbone
        self.head = head
        self.joint_num = joint_num
        self.fused_head = fused_head
        self.deferred = bool(deferred and fused_head)
        # training: deconv blocks 2 and 3 through K9 / K10 (row N1); default = wherever the fused head is used
        self.fused_deconv = fused_head if fused_deconv is None else bool(fused_deconv and fused_head)
        self.criterion = JointLocationLoss()

    def forward(self, input_img, target=None):
        if target is not None and self.fused_head:
            feat = self.head.features(self.backbone(input_img), fused_training=self.fused_deconv)
            fl = self.head.final_layer
            return fused_head_integral_l1_loss(feat, fl.weight, fl.bias, target["coord"], target["vis"], target["have_depth"])
        if target is None and self.deferred:
            fl = self.head.final_layer
            # features(): at inference (eval, no autograd) deconv blocks 2 and 3 run as K9, exactly as in predict()
            return DeferredHeatmap(self.head.features(self.backbone(input_img), fused_training=self.fused_deconv), fl.weight, fl.bias, self.joint_num)
        heatmap = self.head(self.backbone(input_img))
        if target is None:
            return heatmap                             # reference contract, model.py:99-103
        return self.criterion(heatmap, target["coord"], target["vis"], target["have_depth"])

    def predict(self, input_img, flip_pairs=None):
        """Inference: (B, J, 3) voxel coordinates, i.e. main/test.py:62-65 without the full-heat-map gather.  With ``flip_pairs``
        (cfg.flip_test, test.py:67-76) the mirrored image goes through the network as well and the two results are merged on the
        device by one K6 launch."""
        coords, width = self._coords(input_img)
        if flip_pairs is None:
            return coords
        flipped, _ = self._coords(torch.flip(input_img, dims=(3,)))
        return flip_merge(coords, flipped, width, flip_pairs)

    def _coords(self, input_img):
        if self.fused_head and not torch.is_grad_enabled():
            feat = self.head.features(self.backbone(input_img))
            fl = self.head.final_layer
            return fused_head_soft_argmax(feat, fl.weight, fl.bias, self.joint_num), feat.shape[3]
        heatmap = self.forward(input_img)
        return soft_argmax(heatmap, self.joint_num), heatmap.shape[3]


class GraphedPredict:
    """``ResPoseNet.predict`` for one fixed batch shape replayed as ONE CUDA graph: at the reference's ``test_batch_size = 4``
    (main/config.py:44) the ~170 launches of the backbone pace the host, not the device -- the graph takes the host out of the test loop
    (main/test.py:53-65).  The network must be in eval mode; parameters are read in place (load a checkpoint, then build this).
    ``__call__(img)`` copies the batch into the captured input and returns the captured ``(B, J, 3)`` output tensor (valid until the
    next call)."""

    def __init__(self, net, example_img, flip_pairs=None, autocast_dtype=None, warmup=3):
        if net.training:
            raise ValueError("GraphedPredict needs net.eval(): BatchNorm must use its running statistics")
        self.net, self.flip_pairs, self.autocast_dtype = net, flip_pairs, autocast_dtype
        dev = example_img.device
        self.static_in = example_img.detach().clone()
        self.stream = torch.cuda.Stream(dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):         # cuDNN algorithm choice, K9's parameter preparation, workspace allocation: all before capture
                self._run()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_out = self._run()
        torch.cuda.current_stream(dev).wait_stream(self.stream)

    def _run(self):
        with torch.no_grad():
            if self.autocast_dtype is not None:
                with torch.autocast("cuda", dtype=self.autocast_dtype):
                    return self.net.predict(self.static_in, self.flip_pairs)
            return self.net.predict(self.static_in, self.flip_pairs)

    def __call__(self, img):
        if img.shape != self.static_in.shape:
            raise ValueError("GraphedPredict was captured for %s, got %s" % (tuple(self.static_in.shape), tuple(img.shape)))
        self.static_in.copy_(img)
        self.graph.replay()
        return self.static_out


def get_pose_net(cfg, is_train, joint_num, fused_head=False, deferred=False, fused_deconv=None):
    """model.py:105-114.  `cfg` needs `resnet_type` and `depth_dim` (main/config.py:24,28)."""
    backbone = ResNetBackbone(cfg.resnet_type)
    head = HeadNet(joint_num, depth_dim=cfg.depth_dim, inplanes=backbone.out_channels)
    if is_train:
        backbone.init_weights()
        head.init_weights()
    return ResPoseNet(backbone, head, joint_num, fused_head=fused_head, deferred=deferred, fused_deconv=fused_deconv)
