# Builds libzseek_b200/libzseek_b200.so: C host reader + C-ABI launch layer + sm_100a kernels.
NVCC     ?= nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall
CFLAGS   := -std=c11 -O2 -g -fPIC -fvisibility=hidden -Wall -Wextra -Iinclude
SRC      := libzseek_b200/csrc
OUT      := libzseek_b200/libzseek_b200.so

all: $(OUT)

$(SRC)/zsk_cuda.o: $(SRC)/zsk_cuda.cu $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(SRC)/reader.o: $(SRC)/reader.c $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	$(CC) $(CFLAGS) -c $< -o $@

$(OUT): $(SRC)/zsk_cuda.o $(SRC)/reader.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(SRC)/*.o $(OUT)

.PHONY: all oracle clean
