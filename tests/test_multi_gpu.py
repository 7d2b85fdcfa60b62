"""On-hardware multi-GPU correctness: two NCCL ranks each run their contiguous shard of a
tile batch through the CUDA-graph replay (with the all-gather of the flat detection buffer captured in the graph); the
gathered detections must equal a single-GPU run of all the tiles BITWISE.  Skipped unless >= 2 GPUs are visible
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = str(rank), str(world), str(rank)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import test_model_gpu as tm
    tm.DEV = dev
    from oracle.weights import make_tiles
    from wildlifemapper_b200.dist import shard_bounds
    from wildlifemapper_b200.graph import GraphedDetector
    model = tm.build("vit_t", 51)
    tiles = make_tiles(world * B, seed=11)
    lo, hi = shard_bounds(world * B, rank, world)
    det = GraphedDetector(model, batch=B, conf_thr=0.05, nms_score_thr=0.05, iou_thr=0.4, gather=True)
    det(tiles[lo:hi].to(dev))
    torch.cuda.synchronize()
    gathered = [t.cpu() for t in det.gathered_outputs()]
    if rank == 0:
        single = GraphedDetector(model, batch=world * B, conf_thr=0.05, nms_score_thr=0.05, iou_thr=0.4)
        single(tiles.to(dev))
        torch.cuda.synchronize()
        ref = [t.cpu() for t in single.outputs]
        counts, kcnt = ref[1], ref[3]
        ok = torch.equal(gathered[1], counts) and torch.equal(gathered[3], kcnt)
        for b in range(world * B):
            n, k = int(counts[b]), int(kcnt[b])
            ok = ok and torch.equal(gathered[0][b, :n], ref[0][b, :n]) and torch.equal(gathered[2][b, :k], ref[2][b, :k])
        q.put((bool(ok), int(counts.sum()), int(kcnt.sum())))
        del single
    # the captured graphs hold NCCL work: release them before the process group goes away
    del det
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gather_equals_single_gpu_run():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_det, n_keep = q.get(timeout=600)
    for p in procs:
        p.join(60)
    hung = [p for p in procs if p.exitcode is None]
    for p in hung:  # never leave a rank behind on the GPU box
        p.kill()
    assert ok and n_det > 0 and n_keep > 0
    assert not hung and all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def test_model_on_second_device_while_first_is_current():
    """ops must launch on the tensors' device; the C library keeps per-device state (SM count, kernel attributes)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_model_gpu as tm
    from oracle.weights import make_tiles
    from segment_anything.utils.misc import NestedTensor
    tiles = make_tiles(1, seed=2)
    old = tm.DEV
    try:
        tm.DEV = "cuda:0"
        m0 = tm.build("vit_t", 51)
        tm.DEV = "cuda:1"
        m1 = tm.build("vit_t", 51)
    finally:
        tm.DEV = old
    torch.cuda.set_device(0)
    with torch.no_grad():
        o1 = m1(NestedTensor(tiles.to("cuda:1"), None), None)  # device 0 is current
        o0 = m0(NestedTensor(tiles.to("cuda:0"), None), None)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.equal(o0["pred_logits"].cpu(), o1["pred_logits"].cpu())
    assert torch.equal(o0["pred_boxes"].cpu(), o1["pred_boxes"].cpu())
