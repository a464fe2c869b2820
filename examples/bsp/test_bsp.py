import torch
from theanompi_b200 import BSP

if __name__ == "__main__":
    n = max(1, torch.cuda.device_count())
    BSP.sync_type, BSP.exch_strategy = "cdd", "fused"
    rule = BSP()
    rule.model_config = dict(n_epochs=1, max_batches=40, cuda_graph=True, data_kwargs=dict(n_train_files=40 * n, n_val_files=n))
    rule.init(devices=["cuda%d" % i for i in range(n)], modelfile="theanompi_b200.models.alex_net", modelclass="AlexNet")
    rule.wait()
