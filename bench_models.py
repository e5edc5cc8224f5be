"""Networks for bench.py's end-to-end arm that torchvision does not ship.  Random-init weights, synthetic data: only the shape of the work matters.

YOLOv5s (v6.0: width 0.5, depth 0.33), written from the public architecture description -- the model file is not part of the reference
(ppq/samples/Yolo/yolo_5.py:13 expects `Models/yolov5s.v5.onnx`).  Element-wise ops are modules (ppq_b200.executor.Add / Concat) so that the
executor sees them as the reference sees the ONNX graph: Conv -> SiLU fuses like Conv-Sigmoid-Mul (optim/refine.py:210-239), shortcut Adds align
to the larger input, Concats and Upsamples align to their output.  Consistent with bench.yolov5s_tensor_table() (60 convolutions, 7 215 616 weights).
"""
import torch
from torch import nn

from ppq_b200.executor import Add, Concat


class ConvAct(nn.Module):
    def __init__(self, c1, c2, k=1, s=1, p=None):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut=True):
        super().__init__()
        self.cv1, self.cv2 = ConvAct(c, c, 1), ConvAct(c, c, 3)
        self.add = Add() if shortcut else None

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return self.add(x, y) if self.add is not None else y


class C3(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=True):
        super().__init__()
        h = c2 // 2
        self.cv1, self.cv2, self.cv3 = ConvAct(c1, h, 1), ConvAct(c1, h, 1), ConvAct(2 * h, c2, 1)
        self.m = nn.Sequential(*[Bottleneck(h, shortcut) for _ in range(n)])
        self.cat = Concat(1)

    def forward(self, x):
        return self.cv3(self.cat(self.m(self.cv1(x)), self.cv2(x)))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        h = c1 // 2
        self.cv1, self.cv2 = ConvAct(c1, h, 1), ConvAct(4 * h, c2, 1)
        self.m1, self.m2, self.m3 = (nn.MaxPool2d(k, 1, k // 2) for _ in range(3))
        self.cat = Concat(1)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m1(x); y2 = self.m2(y1); y3 = self.m3(y2)
        return self.cv2(self.cat(x, y1, y2, y3))


class YOLOv5s(nn.Module):
    def __init__(self, nc=80):
        super().__init__()
        self.b0, self.b1, self.b2 = ConvAct(3, 32, 6, 2, 2), ConvAct(32, 64, 3, 2), C3(64, 64, 1)
        self.b3, self.b4 = ConvAct(64, 128, 3, 2), C3(128, 128, 2)
        self.b5, self.b6 = ConvAct(128, 256, 3, 2), C3(256, 256, 3)
        self.b7, self.b8, self.b9 = ConvAct(256, 512, 3, 2), C3(512, 512, 1), SPPF(512, 512)
        self.h10, self.up11, self.cat12, self.h13 = ConvAct(512, 256, 1), nn.Upsample(scale_factor=2, mode='nearest'), Concat(1), C3(512, 256, 1, False)
        self.h14, self.up15, self.cat16, self.h17 = ConvAct(256, 128, 1), nn.Upsample(scale_factor=2, mode='nearest'), Concat(1), C3(256, 128, 1, False)
        self.h18, self.cat19, self.h20 = ConvAct(128, 128, 3, 2), Concat(1), C3(256, 256, 1, False)
        self.h21, self.cat22, self.h23 = ConvAct(256, 256, 3, 2), Concat(1), C3(512, 512, 1, False)
        no = 3 * (nc + 5)
        self.det3, self.det4, self.det5 = nn.Conv2d(128, no, 1), nn.Conv2d(256, no, 1), nn.Conv2d(512, no, 1)

    def forward(self, x):
        x = self.b2(self.b1(self.b0(x)))
        p3 = self.b4(self.b3(x))
        p4 = self.b6(self.b5(p3))
        x = self.b9(self.b8(self.b7(p4)))
        h10 = self.h10(x)
        x = self.h13(self.cat12(self.up11(h10), p4))
        h14 = self.h14(x)
        o3 = self.h17(self.cat16(self.up15(h14), p3))
        o4 = self.h20(self.cat19(self.h18(o3), h14))
        o5 = self.h23(self.cat22(self.h21(o4), h10))
        return self.det3(o3), self.det4(o4), self.det5(o5)


def build(name: str) -> nn.Module:
    if name == 'yolov5s': return YOLOv5s()
    if name == 'resnet50':
        import torchvision
        return torchvision.models.resnet50(weights=None)
    raise ValueError(name)
