"""empty stub: recstudio.eval imports torchmetrics.functional as M"""
