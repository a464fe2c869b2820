from theanompi_b200 import BSP

if __name__ == "__main__":
    BSP.sync_type, BSP.exch_strategy = "cdd", "fused"          # reference: rule.exch_strategy='nccl16'
    rule = BSP()
    rule.init(devices=["cuda0", "cuda1"], modelfile="theanompi_b200.models.lasagne_model_zoo.resnet50", modelclass="ResNet50")
    rule.wait()
