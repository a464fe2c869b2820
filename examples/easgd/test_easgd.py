from theanompi_b200 import EASGD

if __name__ == "__main__":
    rule = EASGD()
    # device[0] hosts the center; the workers' elastic exchange is one kernel over NVLink on the center's memory
    rule.init(devices=["cuda0", "cuda1", "cuda2"], modelfile="theanompi_b200.models.cifar10", modelclass="Cifar10_model")
    rule.wait()
