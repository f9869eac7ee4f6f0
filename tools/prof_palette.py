import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from palette_and_histo_gan_b200 import io_utils, dataset_utils
dev = torch.device("cuda:0")
src_np, tgt_np = bench.make_palette_inputs(256, 47)
src, tgt = torch.from_numpy(src_np).to(dev), torch.from_numpy(tgt_np).to(dev)
for _ in range(3):
    s_idx, t_idx, pal = dataset_utils.load_indexed_images(src, tgt, "grayness", check=False)
    oh = io_utils.one_hot(t_idx)
    idx2, oh2 = io_utils.rgba_to_indexed(tgt, pal, with_one_hot=True)
    idx3, rgba3 = io_utils.probabilities_to_indexed(oh, pal)
torch.cuda.synchronize()
print("ok")
