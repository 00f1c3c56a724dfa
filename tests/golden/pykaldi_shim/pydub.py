"""empty stub: lets `import shennong` succeed for golden generation"""
