"""unused by the PLP path (imported by shennong.serializers)"""
