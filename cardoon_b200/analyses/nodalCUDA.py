"""placeholder (written next)"""
def make_nodal_circuit(ckt, termList=None):
    raise NotImplementedError
