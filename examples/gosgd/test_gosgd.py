from theanompi_b200 import GOSGD

if __name__ == "__main__":
    rule = GOSGD()
    rule.init(devices=["cuda0", "cuda1", "cuda2", "cuda3"], modelfile="theanompi_b200.models.cifar10", modelclass="Cifar10_model")
    rule.wait()
