from theanompi_b200 import BSP

if __name__ == "__main__":
    BSP.sync_type, BSP.exch_strategy = "avg", "p2p32"          # Adam model: only parameter averaging (wresnet.py:152-153)
    rule = BSP()
    rule.init(devices=["cuda0", "cuda1"], modelfile="theanompi_b200.models.keras_model_zoo.wresnet", modelclass="Wide_ResNet")
    rule.wait()
