"""import-only"""
