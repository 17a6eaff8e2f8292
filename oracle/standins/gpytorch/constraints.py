import math
import torch
import torch.nn.functional as F


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound)))

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        v = value - self.lower_bound
        return v + torch.log(-torch.expm1(-v))


class GreaterThan(Interval):
    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
