"""Parameter containers with the reference's ResNet-FC layout and initialisation
(models/resnet.py:8-110).  The arithmetic runs inside the fused MLP kernels
(csrc/mlp_f32.cu, csrc/mlp_tc.cu); these modules only own the fp32 master weights so that
checkpoints keep the reference key layout `pts_net.{lin_in,lin_out,blocks.b.fc_{0,1}}.{weight,bias}`."""
import torch
from torch import nn


class ResnetBlockFC(nn.Module):
    """x + fc_1(relu(fc_0(relu(x))))  -- pre-activation residual block, size_in == size_out == size_h."""

    def __init__(self, size_in, size_out=None, size_h=None, beta=0.0):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        if beta > 0 or size_in != size_out:
            raise NotImplementedError("B200 path supports ReLU blocks with equal in/out width only")
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        nn.init.zeros_(self.fc_0.bias)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in", nonlinearity="relu")
        nn.init.zeros_(self.fc_1.bias)
        nn.init.zeros_(self.fc_1.weight)          # reference: residual branch starts at zero (:37)
        self.shortcut = None

    def forward(self, x):
        # reference: models/resnet.py:51-59.  No stand-alone kernel exists for one block: the block only ever runs
        # inside NeRF.forward (models/nerf.py:150), which is the fused kernel here.
        raise NotImplementedError("ResnetBlockFC.forward (reference models/resnet.py:51-59) has no stand-alone B200 "
                                  "kernel: the block is evaluated inside the fused NeRF MLP kernel; call NeRF.forward")


class ResnetFC(nn.Module):
    """lin_in -> n_blocks x ResnetBlockFC -> relu -> lin_out."""

    def __init__(self, d_in, d_out=4, n_blocks=5, d_hidden=128, beta=0.0):
        super().__init__()
        if beta > 0:
            raise NotImplementedError("softplus activations are not supported by the B200 path")
        self.lin_in = nn.Linear(d_in, d_hidden)
        nn.init.zeros_(self.lin_in.bias)
        nn.init.kaiming_normal_(self.lin_in.weight, a=0, mode="fan_in", nonlinearity="relu")
        self.lin_out = nn.Linear(d_hidden, d_out)
        nn.init.zeros_(self.lin_out.bias)
        nn.init.kaiming_normal_(self.lin_out.weight, a=0, mode="fan_in")
        self.n_blocks, self.d_in, self.d_out, self.d_hidden = n_blocks, d_in, d_out, d_hidden
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden, beta=beta) for _ in range(n_blocks)])

    def forward(self, x):
        # reference: models/resnet.py:103-110; only caller is NeRF.forward (models/nerf.py:150) = the fused kernel
        raise NotImplementedError("ResnetFC.forward (reference models/resnet.py:103-110) has no stand-alone B200 "
                                  "kernel: the trunk is evaluated inside the fused NeRF MLP kernel; call NeRF.forward")
