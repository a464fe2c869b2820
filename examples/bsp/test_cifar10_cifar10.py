from theanompi_b200 import BSP

if __name__ == "__main__":
    rule = BSP()
    rule.init(devices=["cuda0", "cuda1"], modelfile="theanompi_b200.models.cifar10", modelclass="Cifar10_model")
    rule.wait()
