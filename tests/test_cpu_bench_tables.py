"""bench.py's workload tables (CPU): the activation / weight shape lists the replay arms are built from."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_resnet50_and_yolov5s_tables():
    import bench
    acts, weights = bench.resnet50_tensor_table()
    assert len(weights) == 54 and sum(bench.numel(s) for s in weights) == 25502912            # torchvision resnet50: 53 convs + fc, weights only
    assert acts[0] == (3, 224, 224) and acts[-1] == (1000,) and len(acts) == 106
    acts, weights = bench.yolov5s_tensor_table()
    # the public YOLOv5s (v6.0): 7.2 M parameters, of which 7 215 616 are convolution weights (60 convolutions incl. the 3 detection heads)
    assert len(weights) == 60 and sum(bench.numel(s) for s in weights) == 7215616
    assert acts[0] == (3, 640, 640) and acts[1] == (32, 320, 320) and acts[-3:] == [(255, 80, 80), (255, 40, 40), (255, 20, 20)]
    half = bench.yolov5s_tensor_table(320)[0]
    assert all(h[0] == f[0] and h[1] * 2 == f[1] for h, f in zip(half, acts))                # shapes scale with the input size


def test_reference_and_gpu_arm_share_one_config():
    import argparse

    import bench
    args = argparse.Namespace(workload='resnet50', batch=32)
    acts, weights = bench.resnet50_tensor_table()
    cfg = bench.workload_config(args, acts, weights, 16)
    assert cfg['batch'] == 32 and cfg['samples_per_gpu_per_step'] == 512 and cfg['observed_tensors'] == 106


def test_yolov5s_module_table_executor_and_cpu_port_agree():
    """bench_models.YOLOv5s (the e2e arm), bench.yolov5s_tensor_table() (the replay arm), ppq_b200.executor's observed set and the CPU port's observed
    set describe the same network: same 60 convolution weights, same 83 observed tensors."""
    import torch

    import bench
    import bench_models
    from oracle.cpu_pipeline import CpuPipeline
    from ppq_b200.executor import TorchExecutor
    net = bench_models.YOLOv5s()
    acts, weights = bench.yolov5s_tensor_table()
    assert sorted(tuple(c.weight.shape) for c in net.modules() if isinstance(c, torch.nn.Conv2d)) == sorted(weights)
    ex = TorchExecutor(bench_models.YOLOv5s(), torch.zeros(1, 3, 64, 64))
    kinds = [op.kind for _, op in ex.quantable_operations()]
    assert kinds.count('Conv') == 60 and kinds.count('Swish') == 57 and kinds.count('Concat') == 13 and kinds.count('Add') == 7 and kinds.count('Resize') == 2
    assert len(ex.observed_configs()) == len(acts) == 83
    ops = dict(ex.quantable_operations())
    assert ops['b0.conv#0'].output_cfg.dominated_by is ops['b0.act#0'].output_cfg          # Conv -> SiLU fusion (Conv - Sigmoid - Mul upstream)
    assert ops['b9.m2#0'].output_cfg.dominated_by is ops['b9.cv1.act#0'].output_cfg        # SPPF max-pools are passive: they share the conv's config
    port = CpuPipeline(bench_models.YOLOv5s(), torch.zeros(1, 3, 64, 64))
    assert len(port.observed()) == 83
