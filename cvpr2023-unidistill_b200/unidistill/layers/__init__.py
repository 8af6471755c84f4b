from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())
