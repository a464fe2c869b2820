# placeholder, replaced below
class Rule(object): pass
class BSP(Rule): pass
class EASGD(Rule): pass
class ASGD(Rule): pass
class GOSGD(Rule): pass
