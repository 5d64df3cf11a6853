import math
import torch
import torch.nn.functional as F
from torch.nn import Parameter
from ..inits import glorot, zeros


class Linear(torch.nn.Module):
    """y = x W^T (+ b); weight [out, in]; 'glorot' = U(-a, a), a = sqrt(6/(in+out))."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None,
                 bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            glorot(self.weight)
        else:  # kaiming_uniform(fan=in, a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in))
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            if self.bias_initializer == "zeros":
                zeros(self.bias)
            else:
                bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
                self.bias.data.uniform_(-bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)
