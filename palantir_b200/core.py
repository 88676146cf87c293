"""placeholder (filled in next)"""
